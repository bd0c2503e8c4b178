#!/usr/bin/env python3
"""Filter time of ONE row band of a multi-GPU run, on one GPU, for several work-unit settings of the symmetric kernel
(SMC_SYM_UNIT="tiles per long unit, tiles per short unit, percent of rows in short units"; unset = the built-in heuristic).
    python tools/sweep_band_units.py [--rows 270] [--width 3840] [--radius 20] auto 1,1,100 2,1,20 3,1,40 ..."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from statmc_b200 import synth  # noqa: E402
from statmc_b200.api import Buffer, Context, Denoiser, f32_factor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=270)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("units", nargs="*", default=["auto"])
    a = ap.parse_args()
    r, W, H = a.radius, a.width, a.rows + 2 * a.radius
    ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    b = synth.moment_buffers(W, H, n=64, config_id=3)
    dev = {k: Buffer.from_array(ctx, b[k]) for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")}
    out = Buffer(ctx, H, W, 3)
    dn = Denoiser(ctx, channels=3, width=W, height=H, radius=r, ds_factor=f32_factor(r / 2.0), n=[dev["n"]], mean=[dev["mean"]],
                  m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]], film=dev["film"], gbufs=[dev["normal"], dev["albedo"]],
                  gbuf_dr_factors=[f32_factor(0.1), f32_factor(0.02)], film_filtered_ptrs=[out], film_filtered=out,
                  denoise_film=True, row_begin=r, row_end=H - r)
    dn.prepass()
    for u in a.units:
        if u == "auto":
            os.environ.pop("SMC_SYM_UNIT", None)
        else:
            os.environ["SMC_SYM_UNIT"] = u
        for _ in range(3):
            dn.filter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dn.filter()
        e1.record()
        torch.cuda.synchronize()
        print("%-10s %-44s filter + gather %.4f ms" % (u, dn.kernel_name, e0.elapsed_time(e1) / 20), flush=True)
    dn.close()


if __name__ == "__main__":
    main()
