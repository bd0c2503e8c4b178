#!/bin/bash
# multi-GPU: parity across processes + scaling bench (run with gpurun --gpus N)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_mgpu.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_mgpu.txt
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("n=%d %.1f Mpix/s %.3f ms/step filter %.3f prepass %.3f e2e %.1f accum %.1f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["e2e"]["value"], d["accum"]["value"]))'
for wl in 4k 8k; do
for mode in peer exchange redundant; do
  echo -n "$wl $mode: "; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --halo $mode 2>gpurun_out/bench_mgpu.err | python -c "$P" || tail -5 gpurun_out/bench_mgpu.err
done; done
echo -n "1 GPU 4k: "; python bench.py --no-cpu-baseline --steps 10 | python -c "$P"
