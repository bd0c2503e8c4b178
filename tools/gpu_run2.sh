#!/bin/bash
mkdir -p gpurun_out
timeout 200 ./build/microbench2 > gpurun_out/microbench2.txt 2>&1; cat gpurun_out/microbench2.txt
echo "== golden"; timeout 300 python tools/make_golden_ref.py gpurun_out/golden > gpurun_out/golden.txt 2>&1; tail -3 gpurun_out/golden.txt
B="python bench.py --no-cpu-baseline --no-accum --no-e2e"
for py in 4 2; do for d in 3 4; do
  echo "== PY=$py D=$d"; SMC_STREAM_PY=$py SMC_STREAM_DEPTH=$d timeout 300 $B --steps 20 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['kernel_ms'], d['fp32']['frac'], d['clocks'], d['config']['kernel'])"
done; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1.csv $B --workload 1080p --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_stream -s 3 -c 1 -f -o gpurun_out/prof_stream_r1 $B --workload 1080p --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepass -s 3 -c 1 -f -o gpurun_out/prof_prepass_r1 $B --workload 4k --steps 1 --warmup 3 > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out
