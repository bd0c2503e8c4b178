#!/usr/bin/env python3
"""The reference's own Estimator (estimator.cpp compiled unmodified, oracle/_ref/libstatmc_ref_estimator.so) running its
Upload / Denoise / Download / Synchronize section (statpath.cpp:406-418, "CUDA time [ns]") on libstatmc_b200 through the
link shim, on the bench workload.  Prints the section's wall time; a reported number, not the bench metric."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from statmc_b200 import synth  # noqa: E402

W, H, r, sd, n = 3840, 2160, 20, 10.0, 64
if len(sys.argv) > 1 and sys.argv[1] == "1080p":
    W, H = 1920, 1080
b = synth.moment_buffers(W, H, n=n, config_id=3)
res = po.ref_estimator_denoise(b, r, sd, reps=8)
ms = res["cuda_time_ns"] / 1e6
print(json.dumps({"what": "reference Estimator::Upload+Denoise+Download+Synchronize on libstatmc_b200 (link shim)",
                  "uploads": "held back until the filter call and moved inside its row-chunked pipeline"
                  if os.environ.get("STATMC_B200_DEFER_UPLOADS", "1") != "0" else "issued by GpuMat::upload, as OpenCV does",
                  "width": W, "height": H, "radius": r, "ms": ms, "mpix_per_s": W * H / ms / 1e3,
                  "planes_registered": res["n_registered"]}))
