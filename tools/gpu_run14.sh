#!/bin/bash
# 4-GPU check of the sharded path: parity test (peer halos + NCCL exchange vs unsharded) and the bench at N = 4 (4K, 8K)
mkdir -p gpurun_out
N=${NGPU:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 400 python -m pytest tests/test_multigpu_gpu.py -q -x --timeout 300 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
for wl in 4k 8k; do
  timeout 300 $TR bench.py --gpus $N --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r14_$wl.err > gpurun_out/bench_n${N}_$wl.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$wl.json").read().strip().splitlines()[-1])
    print("$wl n=%d %.1f Mpix/s %.3f ms/step filter %.3f prepass %.3f e2e %.1f accum %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["e2e"]["value"], d["accum"]["value"]))
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r14_$wl.err").read()[-1500:])
PY
done
