#!/usr/bin/env python3
"""tests/golden/render_veach_mis_16spp.npz: REAL statistics from the reference's own renderer (BASELINE.json configs[0]).

oracle/_ref/pbrt_ref_cpu is the reference's pbrt-v3 + StatPathIntegrator, every source compiled unmodified (oracle/Makefile
target `pbrt`), linked against the host half of integration/opencv_link_shim.cpp and a CPU stand-in for the device half
(oracle/ref_null_device.cpp, which runs the oracle's restatement of the kernels).  This script renders the reference's
scenes/veach-mis/scene-stat.pbrt with scenes/render-denoise.pbrt as the active integrator configuration -- 16 spp in the
4-4-8 schedule, multichannel statistics, denoiseimage, r = 20, sd = 10, G-buffers albedo 0.02 / normal 0.1 -- at a reduced
film size (fixture size; the only edits are the film resolution, the output path, `iterations` 13 -> 3 and the output
regex), and stores the dumped planes of the last iteration.  Runs in the build container only (needs /root/reference).

    python tools/make_golden_render.py                                       # 160 x 90, 16 spp   (configs[0])
    python tools/make_golden_render.py --width 80 --height 45 --iterations 7   # 256 spp: t-table index 509 (configs[1]'s spp)
    python tools/make_golden_render.py --width 80 --height 45 --iterations 11 --config render-denoise-glass-caustics.pbrt
                                                                             # 4096 spp, r 6 / sd 3: table clamp (configs[4]'s)
    python tools/make_golden_render.py --width 80 --height 45 --samples      # 16 spp + the per-pixel radiance sample stream
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from statmc_b200 import pfm  # noqa: E402

REF = "/root/reference"
EXE = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref_cpu")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--height", type=int, default=90)
    ap.add_argument("--iterations", type=int, default=3, help="4 << (iterations - 1) spp in total (expiterations)")
    ap.add_argument("--config", default="render-denoise.pbrt",
                    help="integrator configuration of the reference: render-denoise.pbrt (r 20, sd 10) or "
                         "render-denoise-glass-caustics.pbrt (r 6, sd 3)")
    ap.add_argument("--samples", action="store_true",
                    help="also store the radiance sample stream [S][H][W][3] logged by oracle/_ref/pbrt_ref_cpu_samplelog (the "
                         "same renderer with a forced-include hook in front of the reference's accumulation) -> "
                         "render_veach_mis_<spp>spp_samples.npz")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    spp = 4 << (a.iterations - 1)
    if not a.out:
        a.out = os.path.join(ROOT, "tests", "golden", "render_veach_mis_%dspp%s.npz" % (spp, "_samples" if a.samples else ""))
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "scenes", "veach-mis"))
        os.makedirs(os.path.join(tmp, "out"))
        s = open(os.path.join(REF, "scenes", "veach-mis", "scene-stat.pbrt")).read()
        res = '"integer xresolution" [ 1280 ] "integer yresolution" [ 720 ]'
        name = '"string filename" [ "veach-mis.pfm" ]'
        assert res in s and name in s
        s = s.replace(res, '"integer xresolution" [ %d ] "integer yresolution" [ %d ]' % (a.width, a.height))
        s = s.replace(name, '"string filename" [ "%s/out/veach-mis.pfm" ]' % tmp)
        open(os.path.join(tmp, "scenes", "veach-mis", "scene-stat.pbrt"), "w").write(s)
        c = open(os.path.join(REF, "scenes", a.config)).read()
        it, rx = '"integer  iterations"         [13]', '"string   outputregex"  ["film|film-f"]'
        assert it in c and rx in c
        c = c.replace(it, '"integer  iterations"         [%d]' % a.iterations).replace(rx, '"string   outputregex"  [".*"]')
        import re
        radius = int(re.search(r'"integer\s+filterradius"\s+\[(\d+)\]', c).group(1))
        sd = float(re.search(r'"float\s+filtersd"\s+\[([0-9.]+)\]', c).group(1))
        open(os.path.join(tmp, "scenes", "_active.pbrt"), "w").write(c)  # scene-stat.pbrt: Include "../_active.pbrt"
        lut = os.path.join(tmp, "t005.f32")
        po.t_table(0.005).tofile(lut)
        p = subprocess.run([EXE, "--writeimages", "--nthreads", "8", "scene-stat.pbrt"], text=True, capture_output=True,
                           cwd=os.path.join(tmp, "scenes", "veach-mis"), env=dict(os.environ, STATMC_T_LUT=lut))
        assert p.returncode == 0, p.stdout + p.stderr
        st = os.path.join(tmp, "out", "veach-mis-%d-" % spp)
        rd = lambda k, dt=np.float32: pfm.read(st + k + ".pfm", dt) if dt is not np.float32 else pfm.read(st + k + ".pfm")
        z = {"n": rd("t0-b0-n", np.int32), "mean": rd("t0-b0-mean"), "m2": rd("t0-b0-m2"), "m3": rd("t0-b0-m3"),
             "film_mean": rd("t0-b0-film-mean"), "film_m2": rd("t0-b0-film-m2"), "film": rd("film"),
             "normal": rd("t1-b0-film-mean"), "albedo": rd("t2-b0-film-mean"), "film_f": rd("film-f")}
        assert int(z["n"].min()) == spp and int(z["n"].max()) == spp
        z["config"] = np.array('{"scene": "veach-mis/scene-stat.pbrt + %s", "spp": %d, "radius": %d, "sd": %.1f, '
                               '"normal_sd": 0.1, "albedo_sd": 0.02, "film_f": "reference Estimator flow, kernels = oracle f32"}'
                               % (a.config, spp, radius, sd))
        if a.samples:
            # the logging build renders the same image (the render is deterministic) and writes every radiance sample
            logf = os.path.join(tmp, "samples.f32")
            for k in os.listdir(os.path.join(tmp, "out")):
                os.remove(os.path.join(tmp, "out", k))
            env = dict(os.environ, STATMC_T_LUT=lut, STATMC_SAMPLE_LOG=logf, STATMC_SAMPLE_LOG_W=str(a.width),
                       STATMC_SAMPLE_LOG_H=str(a.height), STATMC_SAMPLE_LOG_S=str(spp))
            p2 = subprocess.run([EXE + "_samplelog", "--writeimages", "--nthreads", "8", "scene-stat.pbrt"], text=True,
                                capture_output=True, cwd=os.path.join(tmp, "scenes", "veach-mis"), env=env)
            assert p2.returncode == 0, p2.stdout + p2.stderr
            for k in ("mean", "m2", "m3", "film-mean", "film-m2"):
                again = rd("t0-b0-" + k)
                assert np.array_equal(again.view(np.uint32), z[k.replace("-", "_")].view(np.uint32)), k
            z["samples"] = np.fromfile(logf, np.float32).reshape(spp, a.height, a.width, 3)
            for k in ("film", "normal", "albedo", "film_f"):  # the denoiser fixtures carry those
                del z[k]
        np.savez_compressed(a.out, **z)
        print(a.out, os.path.getsize(a.out), "bytes;", p.stdout.count("Iteration:"), "iterations")


if __name__ == "__main__":
    main()
