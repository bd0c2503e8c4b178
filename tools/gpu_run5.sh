#!/bin/bash
# A/B: ring refill policy; pipelined e2e; parity tests
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.txt
B="python bench.py --no-cpu-baseline --no-accum --steps 20 --warmup 3"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%.1f Mpix/s filter %.3f ms prepass %.3f ms e2e %.1f Mpix/s (%.2f ms) %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["kernel"]))'
for v in "" refill0; do for wl in 4k 1080p 720p; do
  echo -n "variant=$v $wl: "; SMC_LIB_VARIANT=$v timeout 300 $B --workload $wl 2>&1 | python -c "$P"
done; done
echo -n "8k: "; timeout 600 $B --workload 8k --steps 5 2>&1 | python -c "$P"
