#!/bin/bash
# full GPU suite (no -x: every failure is listed) + the default bench line, after the test / bench edits
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_full.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_full.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_4k_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_4k_r1f.json').read().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['accum']['value'], d['accum'].get('cpu_baseline'), d['cpu_baseline'])"; tail -3 gpurun_out/bench_r1f.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_4k_r1f.json 2>>gpurun_out/bench_r1f.err; tail -c 600 gpurun_out/bench_ref_4k_r1f.json
