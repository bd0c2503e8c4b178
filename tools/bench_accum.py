#!/usr/bin/env python3
"""Stage-1 throughput: smc_accumulate (RGB, Box-Cox transform, M3) for several batch sizes and sample distributions.

    python tools/bench_accum.py [--dist uniform|heavy|real|all] [S ...]        (default S = 4 16 64 256 2048)

Distributions (generated on the device):
  uniform  U(0.01, 4): no zeros, no spikes -- the easy case (never leaves the kernel's fast path)
  heavy    the glass-caustics stand-in of statmc_b200/synth.py (BASELINE configs[4]): Gamma(k = 0.25) radiance with a x1000
           caustic spike with probability 1/512
  real     the distribution of the logged veach-mis sample stream (tests/golden/render_veach_mis_16spp_samples.npz):
           7 % exact zeros, log-normal body, light hits above 1000 with probability 1e-3
Prints Gsamples/s, achieved HBM GB/s on the algorithmic bytes (12 B/sample + 128 B/pixel per launch) and the share of
(pixel, sample) updates that took the scalar IEEE fallback.  The pixel count shrinks with S so that a batch stays <= 12 GB.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from statmc_b200.api import Context, MomentState  # noqa: E402


def samples(dist, S, H, W):
    g = torch.Generator(device="cuda").manual_seed(1234 + S)
    shape = (S, H, W, 3)
    if dist == "uniform":
        return torch.empty(shape, dtype=torch.float32, device="cuda").uniform_(0.01, 4.0, generator=g)
    if dist == "heavy":
        k = 0.25
        x = torch._standard_gamma(torch.full(shape, k, dtype=torch.float32, device="cuda")) * (1.0 / k)
        spike = torch.rand((S, H, W, 1), device="cuda", generator=g) < (1.0 / 512.0)
        return torch.where(spike, x * 1000.0, x).contiguous()
    if dist == "real":
        x = torch.exp(torch.randn(shape, device="cuda", generator=g) * 1.5 - 1.0)
        u = torch.rand((S, H, W, 1), device="cuda", generator=g)
        x = torch.where(u < 0.07, torch.zeros_like(x), x)
        x = torch.where(u > 0.999, x * 3000.0 + 1000.0, x)
        return x.contiguous()
    raise ValueError(dist)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dist", default="all")
    ap.add_argument("--json", default="")
    ap.add_argument("S", nargs="*", type=int)
    a = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    W = 3840
    out = []
    for dist in (["uniform", "heavy", "real"] if a.dist == "all" else [a.dist]):
        for S in a.S or [4, 16, 64, 256, 2048]:
            H = int(max(8, min(1080, 12e9 // (S * W * 12))))
            st = MomentState(ctx, W, H, 3, transform=True)
            smp = samples(dist, S, H, W)
            for _ in range(2):
                st.add_samples_dev(smp.data_ptr(), S)
            ctx.accumulate_fallback_samples()
            st.add_samples_dev(smp.data_ptr(), S)
            slow = ctx.accumulate_fallback_samples() / float(S * H * W)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10 if S <= 64 else 3
            e0.record()
            for _ in range(reps):
                st.add_samples_dev(smp.data_ptr(), S)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / reps * 1e-3
            byts = S * H * W * 12 + H * W * 128
            r = {"dist": dist, "S": S, "rows": H, "ms": t * 1e3, "gsamples_per_s": S * H * W / t / 1e9, "gb_per_s": byts / t / 1e9,
                 "hbm_frac": byts / t / 1e9 / peak, "fallback_share": slow}
            out.append(r)
            print("%-8s S=%4d rows=%4d  %8.3f ms  %6.1f Gsamples/s  %5.0f GB/s = %4.1f%% of %.0f   fallback %.2e" % (
                dist, S, H, r["ms"], r["gsamples_per_s"], r["gb_per_s"], 100 * r["hbm_frac"], peak, slow), flush=True)
            del smp, st
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
