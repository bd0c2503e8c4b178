#!/bin/bash
# scalar-statistics streaming kernel: parity tests, then device-resident timings (RGB at all four BASELINE sizes, scalar 4K stream vs generic)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_replay_gpu.py tests/test_configs_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -6
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%-6s %7.1f Mpix/s  step %.3f ms  filter %.3f ms  prepass %.3f ms  fp32 %.3f  %s" % (d["config"]["workload"].split()[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["fp32"]["frac"], d["config"]["kernel"]))'
for wl in 720p 1080p 4k 8k; do
  timeout 300 python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 10 --warmup 3 --workload $wl 2>gpurun_out/ab.err | tee gpurun_out/bench_dev_$wl.json | python -c "$P" || tail -5 gpurun_out/ab.err
done
for k in 2 1; do
  echo -n "scalar kernel=$k: "; timeout 300 python bench.py --channels 1 --kernel $k --steps 5 --warmup 3 --workload 4k 2>gpurun_out/ab.err | python -c "$P" || tail -5 gpurun_out/ab.err
done
