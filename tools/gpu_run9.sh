#!/bin/bash
# 8-GPU: multi-process parity (world 2/4/8) + scaling of the default bench (4k) and the 8k config
mkdir -p gpurun_out
N=${1:-8}
echo "== pytest multigpu"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_mgpu8.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_mgpu8.txt
P='import sys,json; l=sys.stdin.readlines()[-1]; open("gpurun_out/scale_r1.jsonl","a").write(l); d=json.loads(l); print("n=%d %.1f Mpix/s %.3f ms/step filter %.3f prepass %.3f e2e %.1f (%.2f ms) accum %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["accum"]["value"]))'
for wl in 4k 8k; do for n in 8 4 2; do
  [ $n -le $N ] || continue
  echo -n "$wl n=$n: "; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --workload $wl 2>gpurun_out/bench_mgpu.err | python -c "$P" || tail -5 gpurun_out/bench_mgpu.err
done; done
