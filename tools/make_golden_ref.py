#!/usr/bin/env python3
"""Capture outputs of the REFERENCE'S OWN CUDA kernels (oracle/_ref, compiled unmodified from /root/reference) on
seeded synthetic inputs, as small golden fixtures for the CPU-side tests (tests/test_oracle_cpu.py).

Run on the GPU box:   python tools/make_golden_ref.py gpurun_out/golden
then copy gpurun_out/golden/ref_cuda_*.npz into tests/golden/ and commit them.  The fixtures are self-contained:
inputs (n, mean, m2, m3, film, normal, albedo) AND the reference kernels' outputs (mean-corr, discriminator, film-f)
are stored, because numpy's float32 sin/exp differ by an ulp between CPU generations and a regenerated input would
not be bit-identical on another machine.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    dict(name="a", W=56, H=32, n=16, config_id=81, vary_n=False, radius=6, sd=3.0, normal_sd=0.1, albedo_sd=0.02),
    dict(name="b", W=48, H=40, n=64, config_id=82, vary_n=True, radius=9, sd=5.0, normal_sd=0.1, albedo_sd=0.02),
    dict(name="c", W=44, H=26, n=256, config_id=83, vary_n=False, radius=20, sd=10.0, normal_sd=0.1, albedo_sd=0.02),
]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    os.makedirs(out, exist_ok=True)
    from statmc_b200 import synth
    from statmc_b200.api import Context
    from test_reference_cuda_gpu import _ref_denoise
    ctx = Context(0)
    for c in CASES:
        b = synth.moment_buffers(c["W"], c["H"], n=c["n"], config_id=c["config_id"], vary_n=c["vary_n"])
        r = _ref_denoise(ctx, b, c["radius"], c["sd"], normal_sd=c["normal_sd"], albedo_sd=c["albedo_sd"])
        cfg = {k: v for k, v in c.items() if k != "name"}
        np.savez_compressed(os.path.join(out, "ref_cuda_%s.npz" % c["name"]), config=json.dumps(cfg),
                            mean_corr=r["mean_corr"], disc=r["disc"], film_f=r["film_f"],
                            **{"in_" + k: b[k] for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")})
        print("wrote", c["name"], r["film_f"].shape, float(np.abs(r["film_f"]).mean()))
    ctx.close()


if __name__ == "__main__":
    main()
