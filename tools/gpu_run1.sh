#!/bin/bash
# first GPU session: microbench, tests, golden capture, bench (ours + reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 120 ./build/microbench > gpurun_out/microbench.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -s > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest.txt
echo "== golden"; timeout 300 python tools/make_golden_ref.py gpurun_out/golden > gpurun_out/golden.txt 2>&1; tail -3 gpurun_out/golden.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err; echo "bench rc=$?"; cat gpurun_out/bench_4k.json; tail -5 gpurun_out/bench_4k.err
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_4k.json 2> gpurun_out/bench_ref_4k.err; echo "rc=$?"; cat gpurun_out/bench_ref_4k.json; tail -5 gpurun_out/bench_ref_4k.err
