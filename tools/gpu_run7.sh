#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.txt
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_tmp.json 2>gpurun_out/bench_tmp.err; tail -3 gpurun_out/bench_tmp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tmp.json').read().strip().splitlines()[-1])
print("value %.1f filter %.3f ms prepass %.3f ms (%.0f GB/s, %.2f) e2e %.1f (%.2f ms) accum %.1f Gs/s (%.0f GB/s %.2f)" % (d["value"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["roofline_prepass"]["achieved"], d["roofline_prepass"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["accum"]["value"], d["accum"]["roofline"]["achieved"], d["accum"]["roofline"]["frac"]))
PY
