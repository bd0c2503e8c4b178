#!/usr/bin/env python3
"""The reference's own renderer end to end on the GPU box (oracle/_ref/pbrt_ref_b200: pbrt-v3 + StatPathIntegrator compiled
unmodified, on libstatmc_b200 through the link shim): renders the test scene of tests/render_util.py at a BASELINE size and
reports what the reference itself prints per iteration -- "Rendering time [ns]" (host path tracing + StatTile accumulation)
and "CUDA time [ns]" (Estimator::Upload + Denoise + Download + Synchronize, statpath.cpp:406-418).  A report, not the bench
metric.  NOT yet run on a GPU (written after the round's GPU budget was spent).

    python tools/bench_ref_render.py [--width 1280 --height 720 --iterations 3 --radius 20 --sd 10]
"""
import argparse
import json
import os
import re
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import render_util as ru  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--sd", type=float, default=10.0)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        scene, _ = ru.write_scene(tmp, width=a.width, height=a.height, radius=a.radius, sd=a.sd, iterations=a.iterations)
        p = ru.run_pbrt(ru.PBRT_B200, scene, "--warmup", nthreads=os.cpu_count() or 8)
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
        print(json.dumps({"what": "reference renderer on libstatmc_b200 (link shim); first row = its --warmup pass",
                          "width": a.width, "height": a.height, "radius": a.radius, "iterations": rows}))


if __name__ == "__main__":
    main()
