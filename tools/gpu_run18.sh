#!/bin/bash
# round-end verification of the committed tree: smoke, full GPU suite, default bench + reference arm, raw PCIe probe
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r1g.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r1g.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_4k_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_4k_r1g.json').read().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'accum', d['accum']['value'], 'filter ms', d['roofline']['kernel_ms'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_4k_r1g.json 2>>gpurun_out/bench_r1g.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref_4k_r1g.json').read().splitlines()[-1]); print('ref value', d['value'], 'e2e', d['e2e']['value'], 'accum', d['accum']['value'])"
python - <<'PY'
import torch
n = 630374400
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
ho = torch.empty(99532800, dtype=torch.uint8).pin_memory(); do = torch.empty(99532800, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for both in (False, True):
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("raw pinned H2D of 630 MB%s: %.2f ms = %.1f GB/s" % (" with a concurrent 100 MB D2H" if both else "", ms, n / ms / 1e6))
PY
