#!/bin/bash
# replay CLI tests + fp32 issue-rate microbenchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_replay_gpu.py tests/test_cpp_host_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -15
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb3 tools/microbench3.cu && timeout 120 /tmp/mb3 | tee gpurun_out/microbench3.txt
