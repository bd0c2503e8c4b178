#!/bin/bash
# packed accumulation (two FFMA2 per tap) in the per-warp streaming filter: parity subset + timing at every size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_denoiser_gpu.py tests/test_reference_cuda_gpu.py tests/test_configs_gpu.py tests/test_replay_gpu.py tests/test_reference_estimator_gpu.py -q --timeout 300 -k "not accumulate" 2>&1 | tail -6
for wl in 4k 1080p 720p; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline --no-accum --steps 10 --warmup 3 2>gpurun_out/r17.err | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$wl value %.1f Mpix/s filter %.3f ms e2e %.1f fp32 frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['fp32']['frac']))" || tail -3 gpurun_out/r17.err
done
timeout 200 python bench.py --channels 1 --steps 5 --warmup 3 2>>gpurun_out/r17.err | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('scalar 4k value %.1f Mpix/s filter %.3f ms' % (d['value'], d['roofline']['kernel_ms']))"
