"""Helpers shared by the parity tests."""
import numpy as np


def rel_mad(a, b):
    """relative mean absolute difference: mean|a-b| / mean|b| (the denoiser parity metric, BASELINE.json)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.mean(np.abs(a - b)) / max(np.mean(np.abs(b)), 1e-300))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def moment_rel_err(a, ref, n, sigma_scale):
    """scale-aware relative error for moments (SURVEY.md section 7 'm3 cancellation'):
    max |a - ref| / (|ref| + eps_scale), eps_scale = sigma_scale (e.g. n*sigma^k for the k-th moment)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(a - ref) / (np.abs(ref) + sigma_scale)))


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def bits_equal_nan(a, b):
    """bit equality, except that any NaN matches any NaN (CPU and GPU produce different NaN payloads / signs)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(both_nan | (a.view(np.uint32) == b.view(np.uint32))))


def small_buffers(W=96, H=64, n=32, seed=11, vary_n=False):
    from statmc_b200 import synth
    return synth.moment_buffers(W, H, n=n, config_id=seed, vary_n=vary_n)


def accum_golden():
    """tests/golden/ref_accum_*.npz: the reference's own StatTile accumulation (estimator.h compiled unmodified,
    tools/make_golden_accum.py) on seeded sample batches.  -> list of (name, config, npz)."""
    import glob
    import json
    import os
    out = []
    for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_accum_*.npz"))):
        z = np.load(p)
        out.append((os.path.basename(p)[:-4], json.loads(str(z["config"])), z))
    return out


PLANES = ("mean", "m2", "m3", "film_mean", "film_m2")


def accum_scale(state):
    """per-pixel scale of the k-th central moment sum, for the scale-aware relative error of the accumulation
    parity criterion: sigma for the means, n*sigma^2 for M2, n*sigma^3 for M3 (sigma from the state's
    own M2; floor 1e-30)."""
    n = np.maximum(state["n"].astype(np.float64), 2).reshape(state["n"].shape + (1,))
    m2 = state["m2"].astype(np.float64).reshape(n.shape[:2] + (-1,))
    fm2 = state["film_m2"].astype(np.float64).reshape(n.shape[:2] + (-1,))
    sigma = np.sqrt(np.maximum(m2, 1e-30) / n)
    fsigma = np.sqrt(np.maximum(fm2, 1e-30) / n)
    # states whose M2 is not tracked (M1 configurations) fall back to |mean|
    mean = np.abs(state["mean"].astype(np.float64)).reshape(m2.shape)
    sigma = np.where(m2 > 0, sigma, np.maximum(mean, 1e-30))
    fsigma = np.where(fm2 > 0, fsigma, np.maximum(mean, 1e-30))
    return {"mean": sigma, "m2": n * sigma ** 2, "m3": n * sigma ** 3, "film_mean": fsigma, "film_m2": n * fsigma ** 2}


def real_sample_fixture():
    """tests/golden/render_veach_mis_16spp_samples.npz: the radiance sample stream of the reference's own renderer on its
    veach-mis scene (16 spp, 80 x 45; tools/make_golden_render.py --samples) and the statistic planes its own accumulation
    code made of it.  -> samples [S][H][W][3], reference state dict (n int64)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "render_veach_mis_16spp_samples.npz"))
    ref = {k: z[k] for k in PLANES}
    ref["n"] = z["n"].astype(np.int64)
    return z["samples"], ref


def one_ulp_input_bound(samples):
    """First-order effect on (mean, M2, M3) of the Box-Cox statistics of moving every TRANSFORMED sample by one float32 ulp
    (float64 arithmetic): what replacing libm's powf(x, .5f) by sqrtf(x) -- 1 ulp apart for 6e-4 of inputs -- can do at most,
    pixel by pixel.  d mean / dx_i = 1/n,  d M2 / dx_i = 2 d_i,  d M3 / dx_i = 3 d_i^2 - 3 M2 / n."""
    x = 2.0 * (np.sqrt(samples.astype(np.float64)) - 1.0)
    ulp = np.spacing(np.abs(x).astype(np.float32)).astype(np.float64)
    n = x.shape[0]
    d = x - x.mean(axis=0)
    m2 = (d * d).sum(axis=0)
    return {"mean": ulp.sum(axis=0) / n, "m2": (2.0 * np.abs(d) * ulp).sum(axis=0),
            "m3": (np.abs(3.0 * d * d - 3.0 * m2 / n) * ulp).sum(axis=0)}
