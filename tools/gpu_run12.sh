#!/bin/bash
# packed-f32x2 accumulate kernel: parity tests, accumulate throughput, ncu full capture of the kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moments_gpu.py tests/test_configs_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>gpurun_out/r12.err | tee gpurun_out/bench_r12.json | python -c 'import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("value", d["value"], "accum", d["accum"])' || tail -5 gpurun_out/r12.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate -s 2 -c 1 -f -o gpurun_out/prof_accum_r1d python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
