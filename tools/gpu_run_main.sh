#!/bin/bash
# Main GPU session of a round: smoke, GPU parity tests, bench (ours + reference arm), ncu launch list + full captures.
# Everything lands in gpurun_out/; summaries are copied to profiles/ by tools/summarize_profiles.py afterwards.
R=${ROUND_TAG:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "nproc=$(nproc)" >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_4k_$R.json 2> gpurun_out/bench_4k.err; echo "bench rc=$?"; cat gpurun_out/bench_4k_$R.json; tail -5 gpurun_out/bench_4k.err
for wl in 720p 1080p; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-accum --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$R.json 2>> gpurun_out/bench_4k.err; python -c "import json;d=json.loads(open('gpurun_out/bench_${wl}_$R.json').read().splitlines()[-1]);print('$wl',round(d['value'],1),'Mpix/s e2e',round(d['e2e']['value'],1))"; done
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_4k_$R.json 2> gpurun_out/bench_ref_4k.err; echo "rc=$?"; cat gpurun_out/bench_ref_4k_$R.json; tail -5 gpurun_out/bench_ref_4k.err
B="python bench.py --no-cpu-baseline --no-e2e"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$R.csv $B --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_ -s 3 -c 1 -f -o gpurun_out/prof_filter_$R $B --no-accum --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepass -s 3 -c 1 -f -o gpurun_out/prof_prepass_$R $B --no-accum --steps 1 --warmup 3 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate -s 2 -c 1 -f -o gpurun_out/prof_accum_$R $B --steps 1 --warmup 3 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
ls -la gpurun_out
