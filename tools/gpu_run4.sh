#!/bin/bash
B="python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 20 --warmup 3"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%.1f Mpix/s filter %.3f ms  %s" % (d["value"], d["roofline"]["kernel_ms"], d["config"]["kernel"]))'
run() { echo -n "variant=$1 gbufs=$2 PY=$3: "; SMC_LIB_VARIANT=$1 SMC_STREAM_PY=$3 timeout 300 $B --gbufs $2 2>&1 | python -c "$P"; }
run "" 2 4
run nomufu 2 4
run nomember 2 4
run nomm 2 4
run "" 1 4
run "" 0 4
run nomm 0 4
run "" 0 2
