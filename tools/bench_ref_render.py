#!/usr/bin/env python3
"""The reference's own renderer end to end on the GPU box, at BASELINE configs[0] size, timed by the reference's own timer.

Two links of the SAME renderer (pbrt-v3 + StatPathIntegrator, every source compiled unmodified by oracle/Makefile):
  oracle/_ref/pbrt_ref_b200        Estimator::Upload / Denoise / Download on libstatmc_b200 through the link shim
  oracle/_ref/pbrt_ref_refkernels  same buffers and copies, but stat_denoiser::filter<T> runs the REFERENCE'S kernels
                                   (stat_denoiser.cu compiled unmodified for sm_100a)
Both render the scene of tests/render_util.py (the reference's scenes need its checkout, which does not travel to the GPU box)
at 1280 x 720, 16 spp in the reference's 4-4-8 schedule with scenes/render-denoise.pbrt's parameters, and print what the
reference prints per iteration: "Rendering time [ns]" (host path tracing + StatTile accumulation) and "CUDA time [ns]"
(Upload + Denoise + Download + Synchronize, statpath.cpp:406-418).  The two renders consume the same random numbers, so their
statistic planes are identical and their `film-f` dumps differ by the kernels only: full-size real-data parity against the
reference's kernels (relative MAD, max abs) comes out of the same run.

    python tools/bench_ref_render.py [--width 1280 --height 720 --iterations 3 --radius 20 --sd 10] [--out file.json]
"""
import argparse
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402
from statmc_b200 import pfm  # noqa: E402

REFK = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref_refkernels")


def run(exe, a, tmp, tag):
    d = os.path.join(tmp, tag)
    os.makedirs(d)
    scene, stem = ru.write_scene(d, width=a.width, height=a.height, radius=a.radius, sd=a.sd, iterations=a.iterations,
                                 outputregex="film|film-f")
    p = ru.run_pbrt(exe, scene, "--warmup", "--writeimages", nthreads=os.cpu_count() or 8)
    rows, cur = [], {}
    for line in p.stdout.splitlines():
        m = re.match(r"(Iteration|SPP|Rendering time \[ns\]|CUDA time \[ns\]): (\d+)", line)
        if not m:
            continue
        cur[m.group(1)] = int(m.group(2))
        if m.group(1).startswith("CUDA"):
            rows.append(cur)
            cur = {}
    px = a.width * a.height
    for r in rows:
        r["denoise_mpix_per_s"] = px / (r["CUDA time [ns]"] * 1e-9) / 1e6
    spp = 4 << (a.iterations - 1)
    return rows, pfm.read("%s-%d-film-f.pfm" % (stem, spp)), pfm.read("%s-%d-film.pfm" % (stem, spp))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--sd", type=float, default=10.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = {"what": "the reference's renderer (unmodified) at %dx%d, %d spp; `CUDA time [ns]` = its own timer around "
                   "Upload(); Denoise(); Download(); Synchronize() (statpath.cpp:406-418); first row = its --warmup pass"
                   % (a.width, a.height, 4 << (a.iterations - 1)),
           "width": a.width, "height": a.height, "radius": a.radius, "sd": a.sd, "host_cores": os.cpu_count()}
    with tempfile.TemporaryDirectory() as tmp:
        rows, ff, film = run(ru.PBRT_B200, a, tmp, "ours")
        res["on_libstatmc_b200"] = rows
        if os.path.exists(REFK):
            rows_r, ff_r, film_r = run(REFK, a, tmp, "refk")
            res["with_reference_kernels"] = rows_r
            same_input = bool(np.array_equal(film.view(np.uint32), film_r.view(np.uint32)))
            d = np.abs(ff.astype(np.float64) - ff_r.astype(np.float64))
            res["film_f_vs_reference_kernels"] = {"same_rendered_film": same_input, "rel_mad": float(d.mean() / np.abs(ff_r).mean()),
                                                  "max_abs": float(d.max()), "pixels": int(a.width * a.height)}
            res["cuda_time_ratio_last_iteration"] = rows_r[-1]["CUDA time [ns]"] / rows[-1]["CUDA time [ns]"]
    s = json.dumps(res)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
