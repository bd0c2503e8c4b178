#!/bin/bash
# one ncu --set full capture of the filter kernel (with source), nothing else
R=${ROUND_TAG:-r1}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-accum"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_ -s 3 -c 1 -f -o gpurun_out/prof_filter_$R $B --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
