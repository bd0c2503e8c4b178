#!/usr/bin/env python3
"""Stage-1 throughput only: smc_accumulate on a 3840 x 1080 band (RGB, transform, M3) for several batch sizes.
    python tools/bench_accum.py [S ...]          """
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from statmc_b200.api import Context, MomentState  # noqa: E402

W, H = 3840, 1080
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
for S in [int(a) for a in sys.argv[1:]] or [4, 16, 64]:
    st = MomentState(ctx, W, H, 3, transform=True)
    smp = torch.empty((S, H, W, 3), dtype=torch.float32, device="cuda").uniform_(0.01, 4.0)
    for _ in range(3):
        st.add_samples_dev(smp.data_ptr(), S)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    for _ in range(reps):
        st.add_samples_dev(smp.data_ptr(), S)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / reps * 1e-3
    byts = S * H * W * 12 + H * W * 128
    print("S=%3d  %.3f ms  %.1f Gsamples/s  %.0f GB/s = %.1f%% of %.0f" % (
        S, t * 1e3, S * H * W / t / 1e9, byts / t / 1e9,
        100 * byts / t / 1e9 / peak, peak), flush=True)
    del smp, st
