#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 20 --warmup 3"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%.1f Mpix/s filter %.3f ms  %s clocks=%s" % (d["value"], d["roofline"]["kernel_ms"], d["config"]["kernel"], d["clocks"]["sm_mhz"]))'
run() { echo -n "variant=$1 PY=$2 D=$3: "; SMC_LIB_VARIANT=$1 SMC_STREAM_PY=$2 SMC_STREAM_DEPTH=$3 timeout 300 $B 2>&1 | python -c "$P"; }
run "" 4 4
run "" 2 3
run scalar 4 4
run scalar 2 3
run mb3 2 3
run scalar_mb3 2 3
run scalar_mb3 2 2
run scalar_mb4 2 2
run mb4 2 2
run scalar_mb3 4 3
