#!/usr/bin/env python3
"""Capture outputs of the REFERENCE'S OWN accumulation code (src/statistics/estimator.h, compiled unmodified into
oracle/_ref/libstatmc_ref_accum*.so by oracle/Makefile) on seeded sample streams, as small golden fixtures for
tests/test_oracle_cpu.py and tests/test_moments_gpu.py.

Run in the build container (needs /root/reference for the build, no GPU):
    make -C oracle && python tools/make_golden_accum.py tests/golden
Each fixture is self-contained: the sample batches AND the reference's running totals after every batch are stored
(both the build without FMA contraction, which the restatement follows, and the contracted one, for the spread).
`boxcox_probe` holds the reference's boxCox(x, .5f) of the first batch: it is libm's powf, so a test can tell a libm
that rounds differently from a real mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # radiance configuration (statpath.cpp:1042-1046): Box-Cox transform + M3, RGB, the 4-4-8-16 schedule
    dict(name="a", W=40, H=12, C=3, transform=True, max_moment=3, batches=(4, 4, 8, 16), config_id=91, heavy=False),
    # heavy-tailed stream (config 5 analogue), scalar statistics (multichannelstats=false)
    dict(name="b", W=36, H=10, C=1, transform=True, max_moment=3, batches=(4, 4, 8, 16, 32), config_id=92, heavy=True),
    # feature configuration (statpath.cpp:1117-1118): no transform, M1; and M2 (calcprodenstats)
    dict(name="c", W=32, H=8, C=3, transform=False, max_moment=1, batches=(4, 12), config_id=93, heavy=False),
    dict(name="d", W=32, H=8, C=3, transform=False, max_moment=2, batches=(5, 11), config_id=94, heavy=False),
    dict(name="e", W=32, H=8, C=3, transform=True, max_moment=2, batches=(3, 13), config_id=95, heavy=True),
]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    from oracle import pyoracle as po
    from statmc_b200 import synth
    assert po.ref_accum_available() and po.ref_accum_available(fma=True), "run `make -C oracle` first"
    for c in CASES:
        W, H, C = c["W"], c["H"], c["C"]
        sc = synth.scene(W, H, c["config_id"])
        st, stf = po.new_state(H, W, C), po.new_state(H, W, C)
        arrays, first = {}, 0
        for b, S in enumerate(c["batches"]):
            x = synth.sample_stream(W, H, S, config_id=c["config_id"], first_sample=first, heavy_tail=c["heavy"], sc=sc)
            first += S
            x = np.ascontiguousarray(x[..., :C])
            po.ref_accumulate(st, x, transform=c["transform"], max_moment=c["max_moment"])
            po.ref_accumulate(stf, x, transform=c["transform"], max_moment=c["max_moment"], fma=True)
            arrays["samples_%d" % b] = x
            for k, v in st.items():
                arrays["ref_%d_%s" % (b, k)] = v.copy()
            for k, v in stf.items():
                arrays["reffma_%d_%s" % (b, k)] = v.copy()
        x0 = arrays["samples_0"].ravel()
        arrays["boxcox_probe"] = np.array([po.ref_box_cox(float(v)) for v in x0[:4096]], dtype=np.float32)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items() if k != "name"}
        np.savez_compressed(os.path.join(out, "ref_accum_%s.npz" % c["name"]), config=json.dumps(cfg), **arrays)
        print("wrote ref_accum_%s" % c["name"], {k: float(np.abs(v).mean()) for k, v in st.items() if k != "n"})


if __name__ == "__main__":
    main()
