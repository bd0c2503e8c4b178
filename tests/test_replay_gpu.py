"""build/smc_denoise: the reference's `pbrt --denoise` flow (StatPathIntegrator::Denoise<T>, statpath.cpp:455-550) over PFM
statistic dumps, run as a user would and checked against the oracle on the planes it wrote back."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import pfm, synth
from util import bits_equal, rel_mad

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "smc_denoise")


def _run(*args):
    assert os.path.exists(EXE), "build/smc_denoise missing: run `python __graft_entry__.py`"
    p = subprocess.run([EXE, *[str(a) for a in args]], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p


def test_replay_rgb_default(tmp_path):
    # scenes/denoise.pbrt: RGB radiance statistics (t0) + normal (t1) + albedo (t2) G-buffers, denoiseimage, two iterations
    W, H, r, sd = 150, 70, 8, 4.0
    stem, out = str(tmp_path / "veach"), str(tmp_path / "res")
    want = {}
    for spp, seed in ((4, 21), (8, 22)):
        b = synth.moment_buffers(W, H, n=spp, config_id=seed)
        pfm.write_dump(stem, spp, {"film": b["film"], "t0-b0-n": b["n"], "t0-b0-mean": b["mean"], "t0-b0-m2": b["m2"],
                                   "t0-b0-m3": b["m3"], "t1-b0-film-mean": b["normal"], "t2-b0-film-mean": b["albedo"],
                                   "film-f": np.zeros((H, W, 3), np.float32),       # a previous run's output: not an input
                                   "t9-b0-mean": np.zeros((H, W, 3), np.float32)})  # no such statistic type: skipped, not UB
        want[spp] = po.denoise(b, radius=r, sd=sd, precision="f64", want_aux=True)
    args = ["--stem", stem, "--outstem", out, "--pixelsamples", 4, "--iterations", 2, "--denoiseimage", "true", "--filtersd", sd,
            "--filterradius", r, "--filterbuffers", "albedo,normal", "--filterbuffersds", "0.02,0.1", "--trackedbounces", 0,
            "--outputregex", "film-f|t0-b0-(mean-corr|discriminator|n)"]
    p = _run(*args, "--warmup")
    lines = p.stdout.splitlines()
    assert lines[0] == "==== Warm-Up Start ====" and "==== Warm-Up End ====" in lines
    assert [l for l in lines if l.startswith("Iteration: ")] == ["Iteration: 1", "Iteration: 1", "Iteration: 2"]
    assert sum(l.startswith("CUDA time [ns]: ") for l in lines) == 3 and sum(l.startswith("I/O time [ns]: ") for l in lines) == 3
    assert "skipping veach-4-t9-b0-mean.pfm" in p.stderr
    got = {}
    for spp in (4, 8):
        f = pfm.read("%s-%d-film-f.pfm" % (out, spp))
        got[spp] = f
        assert rel_mad(f, want[spp]["film_f"]) <= 1e-4
        assert bits_equal(pfm.read("%s-%d-t0-b0-mean-corr.pfm" % (out, spp)), want[spp]["mean_corr"])
        assert bits_equal(pfm.read("%s-%d-t0-b0-discriminator.pfm" % (out, spp)), want[spp]["disc"])
        assert np.all(pfm.read("%s-%d-t0-b0-n.pfm" % (out, spp), np.int32) == spp)
    # the pipelined host path (Upload + Denoise + Download as one chunked call) runs the same kernels over row chunks: the
    # symmetric filter then cuts its sums differently (same weights, same decisions, other order of summation)
    out2 = str(tmp_path / "res2")
    _run(*[out2 if a == out else a for a in args], "--pipelined")
    for spp in (4, 8):
        assert rel_mad(pfm.read("%s-%d-film-f.pfm" % (out2, spp)), got[spp]) <= 1e-6


def test_replay_scalar_acrr_smis(tmp_path):
    # multichannelstats=false + acrr + smis, 2 tracked bounces: scalar radiance per bounce (t0-b0, t0-b1), MIS win rates
    # (t1-b*, t2-b*), normal (t3) and albedo (t4) G-buffers; one float-group launch filters all six images, and its z == 0
    # image also filters the RGB film with the scalar gate (stat_denoiser.cu:251-253, 263-265, 271-273)
    W, H, r, sd, spp = 96, 48, 6, 3.0, 16
    stem = str(tmp_path / "box")
    rgb = synth.moment_buffers(W, H, n=spp, config_id=31)
    planes = {"film": rgb["film"], "t3-b0-film-mean": rgb["normal"], "t4-b0-film-mean": rgb["albedo"]}
    images = []
    for t, transform in ((0, True), (1, False), (2, False)):
        for j in (0, 1):
            b = synth.moment_buffers(W, H, n=spp, config_id=40 + 2 * t + j)
            st = {k: np.ascontiguousarray(b[k][:, :, (t + j) % 3]) for k in ("mean", "m2", "m3")}
            value = np.ascontiguousarray(b["film"][:, :, 0]) if transform else st["mean"]  # untransformed types: film-mean IS mean
            pre = "t%d-b%d-" % (t, j)
            planes.update({pre + "n": b["n"], pre + "mean": st["mean"], pre + "m2": st["m2"], pre + "m3": st["m3"]})
            if transform:
                planes[pre + "film-mean"] = value
            images.append((pre, b["n"], st, value))
    pfm.write_dump(stem, spp, planes)
    _run("--stem", stem, "--pixelsamples", spp, "--iterations", 1, "--multichannelstats", "false", "--denoiseimage", "true",
         "--acrr", "true", "--smis", "true", "--trackedbounces", 2, "--filtersd", sd, "--filterradius", r, "--filterbuffers",
         "normal,albedo", "--filterbuffersds", "0.1,0.02", "--outputregex", "film-f|t[0-2]-b[01]-film-mean-f")
    gb, fac, dsf = [rgb["normal"], rgb["albedo"]], [-0.5 / 0.1 ** 2, -0.5 / 0.02 ** 2], -0.5 / (sd * sd)
    for k, (pre, n, st, value) in enumerate(images):
        mc, dc = po.prepass(n, st["mean"], st["m2"], st["m3"])
        ref = po.filter(value, gb, fac, r, dsf, mean_corr=mc, disc=dc, precision="f64")
        assert rel_mad(pfm.read("%s-%d-%sfilm-mean-f.pfm" % (stem, spp, pre)), ref) <= 1e-4, pre
        if k == 0:
            ref_film = po.filter(rgb["film"], gb, fac, r, dsf, mean_corr=mc, disc=dc, precision="f64")
            assert rel_mad(pfm.read("%s-%d-film-f.pfm" % (stem, spp)), ref_film) <= 1e-4
