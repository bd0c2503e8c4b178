"""Deterministic synthetic statistic buffers and sample streams.

The reference's renderer cannot be built offline, so every BASELINE.json config is exercised on synthetic data
of the same shapes: a piecewise-smooth "scene" (soft-edged regions with their own radiance level, albedo, normal
and depth) from which either per-sample radiance is drawn (`sample_stream`, for the accumulation stage) or the
moment planes are drawn directly (`moment_buffers`, for the denoiser).  Counter-based Philox RNG, keyed by
0x53744D43 ^ config id, so every rank / test / bench run sees identical numbers.

Layouts are the reference's: H x W x 3 float32 interleaved (CV_32FC3), H x W int32 for n.
"""
from __future__ import annotations

import numpy as np

KEY = 0x53744D43


def _rng(config_id: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[KEY ^ config_id, stream]))


def scene(W: int, H: int, config_id: int = 0, regions: int = 64, row0: int = 0, rows: int | None = None,
          full_H: int | None = None):
    """Ground truth per pixel: dict(mu[H,W,3], albedo[H,W,3], normal[H,W,3], depth[H,W], shape_k[H,W]).

    row0/rows/full_H select a row band of a taller image without generating the whole image (multi-GPU shards
    generate only their band + halo, identically to the same rows of the full image)."""
    full_H = H if full_H is None else full_H
    rows = H if rows is None else rows
    g = _rng(config_id, 0)
    cx = g.uniform(0, W, regions).astype(np.float32)
    cy = g.uniform(0, full_H, regions).astype(np.float32)
    rad = g.uniform(0.08, 0.35, regions).astype(np.float32) * np.float32(max(W, full_H))
    level = np.exp(g.uniform(np.log(0.01), np.log(8.0), (regions, 3))).astype(np.float32)
    alb = g.uniform(0.05, 0.95, (regions, 3)).astype(np.float32)
    nrm = g.normal(size=(regions, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dep = g.uniform(1.0, 100.0, regions).astype(np.float32)
    kk = g.choice(np.array([0.25, 1.0, 4.0], dtype=np.float32), regions)

    yy = (np.arange(row0, row0 + rows, dtype=np.float32))[:, None]
    xx = (np.arange(W, dtype=np.float32))[None, :]
    # each pixel belongs to the region with the smallest normalised distance (a power diagram: convex cells)
    best = np.full((rows, W), np.inf, dtype=np.float32)
    owner = np.zeros((rows, W), dtype=np.int32)
    for i in range(regions):
        d = ((xx - cx[i]) ** 2 + (yy - cy[i]) ** 2) / (rad[i] * rad[i])
        m = d < best
        best = np.where(m, d, best)
        owner = np.where(m, np.int32(i), owner)
    # smooth intra-region gradient
    grad = (1.0 + 0.25 * np.sin(xx * np.float32(0.013) + yy * np.float32(0.007))).astype(np.float32)
    mu = level[owner] * grad[..., None]
    tilt = np.stack([np.sin(xx * np.float32(0.002)) + 0 * yy, np.cos(yy * np.float32(0.002)) + 0 * xx,
                     np.ones((rows, W), np.float32)], axis=-1).astype(np.float32) * np.float32(0.05)
    normal = nrm[owner] + tilt
    normal /= np.linalg.norm(normal, axis=-1, keepdims=True)
    return {"mu": mu.astype(np.float32), "albedo": alb[owner].astype(np.float32),
            "normal": normal.astype(np.float32), "depth": dep[owner].astype(np.float32),
            "shape_k": kk[owner].astype(np.float32), "owner": owner}


def moment_buffers(W: int, H: int, n=64, config_id: int = 3, row0: int = 0, rows: int | None = None,
                   full_H: int | None = None, vary_n: bool = False):
    """'moment-synth' tier: the statistic planes a render of `n` spp would have left, drawn directly.

    Returns dict with n (int32), mean, m2, m3 (Box-Cox domain), film (untransformed mean), film_m2, normal, albedo,
    depth.  Rows are generated band-wise reproducibly: the noise of row y depends only on (config_id, y)."""
    full_H = H if full_H is None else full_H
    rows = H if rows is None else rows
    sc = scene(W, H, config_id, row0=row0, rows=rows, full_H=full_H)
    mu, k = sc["mu"], sc["shape_k"][..., None]
    if vary_n:
        table = np.array([2, 3, 16, 64, 256, 512, 513, 1024, 4096], dtype=np.int32)
        nn = table[sc["owner"] % len(table)]
    else:
        nn = np.full((rows, W), int(n), dtype=np.int32)
    nf = nn.astype(np.float32)[..., None]

    def per_row_normal(stream, ch):
        out = np.empty((rows, W, ch), dtype=np.float32)
        # one Philox stream per image row => a band equals the same rows of the full image
        for i in range(rows):
            g = np.random.Generator(np.random.Philox(key=[KEY ^ config_id, (stream << 20) | (row0 + i)]))
            out[i] = g.standard_normal((W, ch), dtype=np.float32)
        return out

    z1, z2, z3, z4 = (per_row_normal(s, 3) for s in (1, 2, 3, 4))
    cv2 = 1.0 / k                       # Gamma(k, 1/k): mean 1, variance 1/k
    film = mu * (1.0 + np.sqrt(cv2 / nf) * z1)
    film = np.maximum(film, 0).astype(np.float32)
    # Box-Cox(.5) domain: x = 2 (sqrt(s) - 1); delta-method moments
    bc_mean = 2.0 * (np.sqrt(mu) - 1.0)
    bc_sd = np.sqrt(mu * cv2) * 1.0     # d/ds 2 sqrt(s) = 1/sqrt(s); sd_x ~ sd_s / sqrt(mu)
    mean = bc_mean + bc_sd / np.sqrt(nf) * z2
    m2 = (bc_sd ** 2) * np.maximum(nf - 1.0 + np.sqrt(2.0 * np.maximum(nf - 1.0, 0)) * z3, 0.05 * nf)
    skew = 0.6 / np.sqrt(k)
    m3 = skew * bc_sd ** 3 * nf * (1.0 + 0.2 * z4)
    film_m2 = (mu * mu * cv2) * nf
    zf = per_row_normal(5, 7) * np.float32(0.01)
    out = {"n": nn, "mean": mean.astype(np.float32), "m2": m2.astype(np.float32), "m3": m3.astype(np.float32),
           "film": film, "film_m2": film_m2.astype(np.float32),
           "normal": (sc["normal"] + zf[..., 0:3]).astype(np.float32),
           "albedo": (sc["albedo"] + zf[..., 3:6]).astype(np.float32),
           "depth": (sc["depth"] + zf[..., 6] * 10).astype(np.float32)}
    return {k_: np.ascontiguousarray(v) for k_, v in out.items()}


def sample_stream(W: int, H: int, nsamples: int, config_id: int = 1, first_sample: int = 0, heavy_tail: bool = False,
                  sc=None):
    """'sampled' tier: per-sample RGB radiance [nsamples, H, W, 3] float32 (sample-major), Gamma(k, mu/k) per channel;
    heavy_tail adds the glass-caustics spikes (p = 1/512 a sample is x1000).  Sample s depends only on
    (config_id, s): batches concatenate to the same stream however they are cut."""
    sc = scene(W, H, config_id) if sc is None else sc
    mu, k = sc["mu"], sc["shape_k"][..., None]
    out = np.empty((nsamples, H, W, 3), dtype=np.float32)
    for s in range(nsamples):
        g = np.random.Generator(np.random.Philox(key=[KEY ^ config_id, (7 << 40) | (first_sample + s)]))
        v = g.standard_gamma(np.broadcast_to(k, mu.shape)).astype(np.float32) * (mu / k)
        if heavy_tail:
            spike = g.random(mu.shape[:2], dtype=np.float32) < np.float32(1.0 / 512.0)
            v = np.where(spike[..., None], v * np.float32(1000.0), v)
        out[s] = v
    return out


def feature_stream(W: int, H: int, nsamples: int, config_id: int = 1, first_sample: int = 0, sc=None):
    """Per-sample normal and albedo features (truth + N(0, 0.01^2) sub-pixel jitter): two [S, H, W, 3] arrays."""
    sc = scene(W, H, config_id) if sc is None else sc
    nrm = np.empty((nsamples, H, W, 3), dtype=np.float32)
    alb = np.empty((nsamples, H, W, 3), dtype=np.float32)
    for s in range(nsamples):
        g = np.random.Generator(np.random.Philox(key=[KEY ^ config_id, (9 << 40) | (first_sample + s)]))
        nrm[s] = sc["normal"] + g.standard_normal((H, W, 3), dtype=np.float32) * np.float32(0.01)
        alb[s] = sc["albedo"] + g.standard_normal((H, W, 3), dtype=np.float32) * np.float32(0.01)
    return nrm, alb
