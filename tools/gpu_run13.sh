#!/bin/bash
# link-level drop-in (reference Estimator on libstatmc_b200): new GPU tests, C++ host tests, quick bench sanity, and the
# reference Estimator's own "CUDA time" section at 4K on our library
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_reference_estimator_gpu.py tests/test_reference_estimator_cpu.py tests/test_cpp_host_gpu.py tests/test_replay_gpu.py -q -x --timeout 300 2>&1 | tail -15
timeout 300 python tools/bench_ref_estimator.py 2>&1 | tail -6 | tee gpurun_out/ref_estimator_4k.txt
timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>gpurun_out/r13.err | tee gpurun_out/bench_r13.json | python -c 'import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("value", d["value"], "e2e", d["e2e"]["value"], "accum", d["accum"]["value"])' || tail -5 gpurun_out/r13.err
