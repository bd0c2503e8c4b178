"""Link-level drop-in on the GPU: the REFERENCE'S OWN pbrt::Estimator (src/statistics/estimator.cpp + buffer.cpp compiled
unmodified, oracle/_ref/libstatmc_ref_estimator.so) allocates its buffers, uploads, denoises and downloads through
integration/opencv_link_shim.cpp, i.e. on libstatmc_b200's C ABI and CUDA kernels.  Results are held to the oracle, to our
own plan API and to the reference's `--denoise` replay flow."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import pfm, synth
from statmc_b200.api import denoise_host
from util import bits_equal, rel_mad

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not po.ref_estimator_available(),
                                 reason="oracle/_ref/libstatmc_ref_estimator.so not built (needs /root/reference)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("W,H,radius,sd,n,vary", [(160, 90, 20, 10.0, 16, False), (301, 57, 6, 3.0, 64, True)])
def test_reference_estimator_rgb_default(ctx, W, H, radius, sd, n, vary):
    # scenes/render-denoise.pbrt: multichannelstats, denoiseimage, filterbuffers normal + albedo
    b = synth.moment_buffers(W, H, n=n, config_id=81, vary_n=vary)
    got = po.ref_estimator_denoise(b, radius, sd)
    # film + film-f + 10 planes for the radiance type + 10 per feature type (estimator.cpp:121-146)
    assert got["n_registered"] == 2 + 3 * 10
    ora = po.denoise(b, radius=radius, sd=sd, precision="f64", want_aux=True)
    assert bits_equal(got["mean_corr"], ora["mean_corr"]) and bits_equal(got["disc"], ora["disc"])
    assert rel_mad(got["film_f"], ora["film_f"]) <= 1e-4
    # t0-b0-film-mean-f shares the host matrix of film-f (estimator.cpp:143-144)
    assert bits_equal(got["film_mean_f"], got["film_f"])
    # and what our own plan API gives on the same planes (the shim's held-back uploads travel through the row-chunked
    # pipeline: the symmetric kernel then sums in another order)
    ours = denoise_host(ctx, b, radius=radius, sd=sd, want_aux=True)
    assert rel_mad(got["film_f"], ours["film_f"]) <= 1e-6


def test_reference_estimator_scalar_acrr_bounces(ctx):
    # multichannelstats=false + acrr, two tracked bounces: ONE filter<float> launch over both images; its image 0 also
    # filters the RGB film with the scalar gate (stat_denoiser.cu:251-253, 263-265, 271-273)
    W, H, r, sd, spp = 120, 50, 7, 3.5, 32
    rgb = synth.moment_buffers(W, H, n=spp, config_id=82)
    per = []
    for j in (0, 1):
        s = synth.moment_buffers(W, H, n=spp, config_id=83 + j)
        per.append({"n": s["n"], **{k: np.ascontiguousarray(s[k][..., j]) for k in ("mean", "m2", "m3")},
                    "film_mean": np.ascontiguousarray(s["film"][..., 1])})
    b = {k: np.stack([p[k] for p in per]) for k in ("n", "mean", "m2", "m3", "film_mean")}
    b.update(film=rgb["film"], normal=rgb["normal"], albedo=rgb["albedo"])
    got = po.ref_estimator_denoise(b, r, sd, acrr=True)
    gb, fac, dsf = [rgb["normal"], rgb["albedo"]], [-0.5 / 0.1 ** 2, -0.5 / 0.02 ** 2], -0.5 / (sd * sd)
    for j, p in enumerate(per):
        mc, dc = po.prepass(p["n"], p["mean"], p["m2"], p["m3"])
        assert bits_equal(got["mean_corr"][j, ..., 0], mc) and bits_equal(got["disc"][j, ..., 0], dc)
        ref = po.filter(p["film_mean"], gb, fac, r, dsf, mean_corr=mc, disc=dc, precision="f64")
        assert rel_mad(got["film_mean_f"][j, ..., 0], ref) <= 1e-4, j
        if j == 0:
            ref_film = po.filter(rgb["film"], gb, fac, r, dsf, mean_corr=mc, disc=dc, precision="f64")
            assert rel_mad(got["film_f"], ref_film) <= 1e-4


def test_reference_estimator_smis_both_cuda_groups(ctx):
    # multichannelstats + denoiseimage + smis, 2 tracked bounces: RGB radiance (t0) in the float3 group, BSDF / light win
    # rates (t1, t2; scalar, untransformed) in the float group.  Estimator::Denoise launches filter<float> first -- its
    # image 0 (t1-b0) filters `film` into `film-f` with the win-rate gate -- and filter<float3> then overwrites film-f with
    # the radiance-gated result (estimator.cpp:434-488).  Both plans stay cached side by side.
    W, H, r, sd, spp, nbm = 110, 52, 6, 3.0, 32, 2
    b = synth.moment_buffers(W, H, n=spp, config_id=86)
    mis = {k: [] for k in ("n", "mean", "m2", "m3")}
    for t in (0, 1):
        for j in range(nbm):
            s = synth.moment_buffers(W, H, n=spp, config_id=87 + 2 * t + j)
            mis["n"].append(s["n"])
            for k in ("mean", "m2", "m3"):
                mis[k].append(np.ascontiguousarray(s[k][..., (t + j) % 3]))
    mis = {k: np.stack(v).reshape(2, nbm, H, W) for k, v in mis.items()}
    for rep in range(2):  # second call: both cached plans are reused
        got = po.ref_estimator_denoise(b, r, sd, mis=mis)
        assert got["n_registered"] == 2 + 10 * (1 + 2 * nbm + 2)
        ora = po.denoise(b, radius=r, sd=sd, precision="f64", want_aux=True)
        assert bits_equal(got["mean_corr"], ora["mean_corr"]) and bits_equal(got["disc"], ora["disc"])
        assert rel_mad(got["film_f"], ora["film_f"]) <= 1e-4  # the RGB launch has the last word on film-f
        gb, fac, dsf = [b["normal"], b["albedo"]], [-0.5 / 0.1 ** 2, -0.5 / 0.02 ** 2], -0.5 / (sd * sd)
        for t in (0, 1):
            for j in range(nbm):
                mc, dc = po.prepass(mis["n"][t, j], mis["mean"][t, j], mis["m2"][t, j], mis["m3"][t, j])
                ref = po.filter(mis["mean"][t, j], gb, fac, r, dsf, mean_corr=mc, disc=dc, precision="f64")
                assert rel_mad(got["mis_f"][t, j], ref) <= 1e-4, (rep, t, j)


def test_reference_dump_feeds_our_replay(ctx, tmp_path):
    # the reference's OutputBufferSelection::Write (buffer.cpp:40-53) dumps every registered plane as PFM through the shim;
    # build/smc_denoise (our `pbrt --denoise` replay) reads that dump back and must reproduce film-f bit for bit
    exe = os.path.join(ROOT, "build", "smc_denoise")
    assert os.path.exists(exe), "build/smc_denoise missing: run `python __graft_entry__.py`"
    W, H, r, sd, spp = 140, 60, 9, 4.5, 8
    b = synth.moment_buffers(W, H, n=spp, config_id=85)
    stem = str(tmp_path / "scene")
    got = po.ref_estimator_denoise(b, r, sd, dump_stem=stem, dump_regex=".*", dump_suffix=str(spp))
    written = sorted(os.listdir(tmp_path))
    assert len(written) == got["n_registered"] and "scene-%d-t0-b0-discriminator.pfm" % spp in written
    assert bits_equal(pfm.read("%s-%d-film-f.pfm" % (stem, spp)), got["film_f"])
    assert bits_equal(pfm.read("%s-%d-t0-b0-m3.pfm" % (stem, spp)), b["m3"])
    assert np.array_equal(pfm.read("%s-%d-t0-b0-n.pfm" % (stem, spp), np.int32), b["n"])
    assert bits_equal(pfm.read("%s-%d-t1-b0-film-mean.pfm" % (stem, spp)), b["normal"])
    out = str(tmp_path / "replayed")
    p = subprocess.run([exe, "--stem", stem, "--outstem", out, "--pixelsamples", str(spp), "--iterations", "1",
                        "--denoiseimage", "true", "--filtersd", str(sd), "--filterradius", str(r), "--filterbuffers",
                        "normal,albedo", "--filterbuffersds", "0.1,0.02", "--trackedbounces", "0", "--outputregex", "film-f"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert bits_equal(pfm.read("%s-%d-film-f.pfm" % (out, spp)), got["film_f"])
