"""REAL statistics from the reference's own renderer on the GPU path.

* the veach-mis fixture (tests/golden/render_veach_mis_16spp.npz, see test_reference_render_cpu.py) through the C ABI;
* oracle/_ref/pbrt_ref_b200 -- the reference's pbrt-v3 + StatPathIntegrator compiled unmodified and linked against
  libstatmc_b200.so through integration/opencv_link_shim.cpp -- rendering a small scene of ours on the host cores with
  every Estimator::Upload / Denoise / Download of its render loop (statpath.cpp:406-418) running on the B200, then its
  `--denoise` replay of the dump it wrote."""
import json
import os

import numpy as np
import pytest

import render_util as ru
from oracle import pyoracle as po
from statmc_b200 import pfm
from statmc_b200.api import denoise_host
from util import bits_equal, max_abs, rel_mad

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# veach-mis from the reference's own renderer: 16 spp (BASELINE configs[0]), 256 spp (t-table index 509; configs[1]'s sample
# count) and 4096 spp with the glass-caustics filter parameters r 6 / sd 3 (table clamp at index 1023; configs[4]'s)
RENDERS = ["render_veach_mis_16spp.npz", "render_veach_mis_256spp.npz", "render_veach_mis_4096spp.npz"]


def _oracle64(b, radius, sd):
    dsf, fac = ru.reference_factors(sd)
    mc, dc = po.prepass(b["n"], b["mean"], b["m2"], b["m3"])
    return po.filter(b["film"], [b["normal"], b["albedo"]], fac, radius, dsf, mean_corr=mc, disc=dc, precision="f64"), mc, dc


@pytest.mark.parametrize("name", RENDERS)
def test_veach_mis_fixture_through_the_c_abi(ctx, name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    cfg = json.loads(str(z["config"]))
    b = {k: z[k] for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")}
    ref, mc, dc = _oracle64(b, cfg["radius"], cfg["sd"])
    for kernel in (2, 1):
        ours = denoise_host(ctx, b, radius=cfg["radius"], sd=cfg["sd"], kernel=kernel, want_aux=True)
        assert bits_equal(ours["mean_corr"], mc) and bits_equal(ours["disc"], dc)
        rm = rel_mad(ours["film_f"], ref)
        print("veach-mis %d spp (%s): relMAD vs f64 transcription %.2e, max-abs %.2e; vs the reference flow's film-f %.2e"
              % (cfg["spp"], ours["kernel"], rm, max_abs(ours["film_f"], ref), rel_mad(ours["film_f"], z["film_f"])))
        assert rm <= 1e-4 and rel_mad(ours["film_f"], z["film_f"]) <= 1e-4


@pytest.mark.skipif(not os.path.exists(ru.PBRT_B200), reason="oracle/_ref/pbrt_ref_b200 not built (needs /root/reference)")
def test_reference_renderer_on_libstatmc_b200(tmp_path):
    scene, stem = ru.write_scene(tmp_path, width=160, height=96, radius=10, sd=5.0)
    p = ru.run_pbrt(ru.PBRT_B200, scene, "--writeimages", nthreads=os.cpu_count() or 8)
    assert [l for l in p.stdout.splitlines() if l.startswith("SPP: ")] == ["SPP: 4", "SPP: 4", "SPP: 8"]
    assert sum(l.startswith("CUDA time [ns]: ") for l in p.stdout.splitlines()) == 3
    first = {}
    for spp in (4, 8, 16):
        b = ru.read_dump(stem, spp)
        assert int(b["n"].min()) == int(b["n"].max()) == spp
        ref, _, _ = _oracle64(b, 10, 5.0)
        assert rel_mad(b["film_f"], ref) <= 1e-4, spp
        first[spp] = b["film_f"]
    for spp in (4, 8, 16):
        os.remove("%s-%d-film-f.pfm" % (stem, spp))
    ru.run_pbrt(ru.PBRT_B200, scene, "--denoise", "--writeimages")
    for spp in (4, 8, 16):
        assert bits_equal(ru.read_dump(stem, spp)["film_f"], first[spp]), spp


REFK = os.path.join(os.path.dirname(ru.PBRT_B200), "pbrt_ref_refkernels")


@pytest.mark.skipif(not (os.path.exists(ru.PBRT_B200) and os.path.exists(REFK)),
                    reason="oracle/_ref/pbrt_ref_b200 / pbrt_ref_refkernels not built (need /root/reference)")
def test_real_render_against_the_reference_kernels(tmp_path):
    # Two links of the reference's renderer render the same frame (same random numbers): one denoises on libstatmc_b200, the
    # other with the reference's own kernels on the same buffers.  Real 16-spp statistics at 640 x 360 (tools/bench_ref_render.py
    # does the same at BASELINE configs[0]'s 1280 x 720 and records the reference's own timer).
    out = {}
    for tag, exe in (("ours", ru.PBRT_B200), ("refk", REFK)):
        d = tmp_path / tag
        d.mkdir()
        scene, stem = ru.write_scene(d, width=640, height=360, radius=20, sd=10.0, outputregex="film|film-f")
        ru.run_pbrt(exe, scene, "--writeimages", nthreads=os.cpu_count() or 8)
        out[tag] = (pfm.read("%s-16-film.pfm" % stem), pfm.read("%s-16-film-f.pfm" % stem))
    assert bits_equal(out["ours"][0], out["refk"][0]), "the two links rendered different frames"
    rm = rel_mad(out["ours"][1], out["refk"][1])
    print("real 640x360 16-spp render: film-f relMAD vs the reference's kernels %.2e, max-abs %.2e" % (rm, max_abs(out["ours"][1], out["refk"][1])))
    assert rm <= 1e-4
