#!/bin/bash
# dynamic tile scheduling: parity, timeline, occupancy/depth variants
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.txt
B="python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 20 --warmup 3"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%.1f Mpix/s filter %.3f ms %s" % (d["value"], d["roofline"]["kernel_ms"], d["config"]["kernel"]))'
run() { echo -n "variant=$1 PY=$2 D=$3 $4: "; SMC_LIB_VARIANT=$1 SMC_STREAM_PY=$2 SMC_STREAM_DEPTH=$3 timeout 300 $B --workload ${4:-4k} 2>&1 | python -c "$P"; }
run "" 2 3
run "" 2 2
run "" 4 4
run "" 4 3
run mb4 2 2
run "" 2 3 1080p
run "" 2 3 720p
SMC_STREAM_TRACE=gpurun_out/trace_4k_dyn.txt python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 3 > /dev/null
