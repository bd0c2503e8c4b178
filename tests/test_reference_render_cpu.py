"""REAL statistics from the reference's own renderer (BASELINE.json configs[0]), no GPU needed.

* tests/golden/render_veach_mis_16spp.npz -- the reference's veach-mis scene, StatPathIntegrator, 16 spp (4-4-8), rendered by
  its own code compiled unmodified (oracle/_ref/pbrt_ref_cpu, tools/make_golden_render.py); `film_f` in it was produced by
  the reference's Estimator::Upload/Denoise/Download flow with the oracle's kernels behind the OpenCV surface.
* a live run of the same binary on a small scene of ours (tests/render_util.py), when the binary is present.
Both pin the numpy front-end of the oracle (which the GPU parity tests use) to the reference's dispatch and routing."""
import json
import os

import numpy as np
import pytest

import render_util as ru
from oracle import pyoracle as po
from util import bits_equal, rel_mad

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# veach-mis from the reference's own renderer: 16 spp (BASELINE configs[0]), 256 spp (t-table index 509; configs[1]'s sample
# count) and 4096 spp with the glass-caustics filter parameters r 6 / sd 3 (table clamp at index 1023; configs[4]'s)
RENDERS = ["render_veach_mis_16spp.npz", "render_veach_mis_256spp.npz", "render_veach_mis_4096spp.npz"]


def _oracle(b, radius, sd, precision):
    dsf, fac = ru.reference_factors(sd)
    mc, dc = po.prepass(b["n"], b["mean"], b["m2"], b["m3"])
    return po.filter(b["film"], [b["normal"], b["albedo"]], fac, radius, dsf, mean_corr=mc, disc=dc, precision=precision)


@pytest.mark.parametrize("name", RENDERS)
def test_veach_mis_fixture_is_what_the_oracle_computes(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    cfg = json.loads(str(z["config"]))
    b = {k: z[k] for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")}
    assert int(b["n"].min()) == int(b["n"].max()) == cfg["spp"] and "%dspp" % cfg["spp"] in name
    # a path-traced image, not a synthetic one: three lights five orders of magnitude apart, black background pixels
    assert float(b["film"].max()) > 100 * float(b["film"].mean()) and int((b["m2"].sum(axis=2) == 0).sum()) > 50
    # Box-Cox statistics: mean of 2 (sqrt(x) - 1) >= -2; the untransformed film-mean is the film up to its colour round trip
    assert float(b["mean"].min()) >= -2.0 and rel_mad(z["film_mean"], b["film"]) < 1e-4  # film: RGB -> XYZ -> RGB in core/film.cpp
    assert bits_equal(_oracle(b, cfg["radius"], cfg["sd"], "f32"), z["film_f"])
    assert rel_mad(_oracle(b, cfg["radius"], cfg["sd"], "f64"), z["film_f"]) < 1e-5


@pytest.mark.skipif(not os.path.exists(ru.PBRT_CPU), reason="oracle/_ref/pbrt_ref_cpu not built (needs /root/reference)")
def test_reference_renderer_runs_and_its_denoise_flow_matches_the_oracle(tmp_path):
    lut = str(tmp_path / "t005.f32")
    po.t_table(0.005).tofile(lut)
    scene, stem = ru.write_scene(tmp_path, width=96, height=64, radius=8, sd=4.0)
    p = ru.run_pbrt(ru.PBRT_CPU, scene, "--writeimages", env={"STATMC_T_LUT": lut})
    assert [l for l in p.stdout.splitlines() if l.startswith("SPP: ")] == ["SPP: 4", "SPP: 4", "SPP: 8"]
    first = {}
    for spp in (4, 8, 16):
        b = ru.read_dump(stem, spp)
        assert int(b["n"].min()) == int(b["n"].max()) == spp
        assert bits_equal(_oracle(b, 8, 4.0, "f32"), b["film_f"]), spp
        first[spp] = b["film_f"]
    # the reference's `--denoise` flow (StatPathIntegrator::Denoise<T>, statpath.cpp:456-550): cv::glob + cv::imread +
    # convertTo + cvtColor of the dump it has just written, then the same Upload / Denoise / Download
    for spp in (4, 8, 16):
        os.remove("%s-%d-film-f.pfm" % (stem, spp))
    ru.run_pbrt(ru.PBRT_CPU, scene, "--denoise", "--writeimages", env={"STATMC_T_LUT": lut})
    for spp in (4, 8, 16):
        assert bits_equal(ru.read_dump(stem, spp)["film_f"], first[spp]), spp


@pytest.mark.skipif(not os.path.exists(ru.PBRT_CPU), reason="oracle/_ref/pbrt_ref_cpu not built (needs /root/reference)")
@pytest.mark.parametrize("mode", ["acrr", "smis"])
def test_reference_renderer_acrr_and_smis_configurations(tmp_path, mode):
    """scenes/acrr.pbrt and scenes/smis.pbrt of the reference: scalar statistics per tracked bounce, one filter<float> launch
    over all of them (no film), and the filtered planes feed the NEXT iteration's Russian roulette / MIS decisions
    (statpath.cpp:306-313) -- the feedback loop runs here with the oracle's kernels in it."""
    from statmc_b200 import pfm
    lut = str(tmp_path / "t005.f32")
    po.t_table(0.005).tofile(lut)
    nb = 5 if mode == "acrr" else 6
    scene, stem = ru.write_scene(tmp_path, width=80, height=48, radius=6, sd=3.0, trackedbounces=nb, multichannelstats=False,
                                 denoiseimage=False, acrr=mode == "acrr", smis=mode == "smis")
    ru.run_pbrt(ru.PBRT_CPU, scene, "--writeimages", env={"STATMC_T_LUT": lut})
    dsf, fac = ru.reference_factors(3.0)
    # statistic types in the order CreateStatPathIntegrator enables them (statpath.cpp:1026-1160)
    # acrr: t0 = radiance per bounce, t1 / t2 = normal / albedo;  smis alone: no radiance statistics at all, t0 / t1 = BSDF /
    # light win rates per bounce, t2 / t3 = normal / albedo
    filtered = [0] if mode == "acrr" else [0, 1]
    i_normal = 1 if mode == "acrr" else 2
    normal = pfm.read("%s-16-t%d-b0-film-mean.pfm" % (stem, i_normal))
    albedo = pfm.read("%s-16-t%d-b0-film-mean.pfm" % (stem, i_normal + 1))
    assert normal.shape == (48, 80, 3) and float(np.abs(normal).max()) <= 1.0 + 1e-6
    checked = 0
    for t in filtered:
        for j in range(nb):
            pl = ru.read_planes(stem, 16, t, j, 1)
            assert pl["n"].shape == (48, 80)
            mc, dc = po.prepass(pl["n"], pl["mean"], pl["m2"], pl["m3"])
            ref = po.filter(pl["film_mean"], [normal, albedo], fac, 6, dsf, mean_corr=mc, disc=dc, precision="f32")
            assert bits_equal(ref, pl["film_mean_f"]), (t, j)
            checked += 1
    assert checked == len(filtered) * nb


@pytest.mark.skipif(not os.path.exists(ru.PBRT_CPU), reason="oracle/_ref/pbrt_ref_cpu not built (needs /root/reference)")
def test_reference_mean_vars_cpu_loop_vs_the_cuda_kernel_semantics(tmp_path):
    """scenes/render-for-proden.pbrt: the estimator-variance planes `film-mean-var` come from the CPU loop the reference ships
    (estimator.cpp:524-568).  For RGB planes that loop multiplies by the reciprocal (OpenCV's Vec3f / float), the CUDA kernel
    it replaced -- and smc_calculate_mean_vars, which follows the kernel -- divides: at most 1 ulp apart."""
    from statmc_b200 import pfm
    lut = str(tmp_path / "t005.f32")
    po.t_table(0.005).tofile(lut)
    scene, stem = ru.write_scene(tmp_path, width=80, height=48, denoiseimage=False, calcprodenstats=True)
    p = ru.run_pbrt(ru.PBRT_CPU, scene, "--writeimages", env={"STATMC_T_LUT": lut})
    assert "CUDA time [ns]: 0" in p.stdout                      # nothing to denoise: runCUDA stays false (statpath.cpp:406)
    for t in (0, 1, 2):                                          # radiance, normal, albedo
        n = pfm.read("%s-16-t%d-b0-n.pfm" % (stem, t), np.int32)
        m2 = pfm.read("%s-16-t%d-b0-film-m2.pfm" % (stem, t))
        var = pfm.read("%s-16-t%d-b0-film-mean-var.pfm" % (stem, t))
        assert bits_equal(po.calculate_mean_vars_cpu_loop(n, m2), var), t
        kernel = po.calculate_mean_vars(n, m2)                   # stat_denoiser.cu:148-159 semantics
        ulp = np.spacing(np.abs(var))
        assert np.all(np.abs(kernel - var) <= ulp) and float(np.mean(kernel != var)) > 0.05
