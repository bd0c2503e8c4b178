#!/bin/bash
# A/B timing of filter-kernel variants: each argument is an environment assignment list ("VAR=x VAR2=y", or "-" for none);
# runs the GPU parity tests of the denoiser once with the first configuration, then the device-resident bench for each.
mkdir -p gpurun_out
WL=${WORKLOADS:-4k}
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%-10s %7.1f Mpix/s  step %.3f ms  filter %.3f ms  prepass %.3f ms  %s" % (d["config"]["workload"].split()[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["config"]["kernel"]))'
if [ -z "$SKIP_TESTS" ]; then
  echo "== pytest ($1)"; env $( [ "$1" = "-" ] || echo $1 ) timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_configs_gpu.py tests/test_reference_cuda_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -4
fi
for cfg in "$@"; do
  for wl in $WL; do
    echo -n "[$cfg] "; env $( [ "$cfg" = "-" ] || echo $cfg ) timeout 300 python bench.py --no-cpu-baseline --no-accum --no-e2e --steps 10 --warmup 3 --workload $wl 2>gpurun_out/ab.err | python -c "$P" || tail -5 gpurun_out/ab.err
  done
done
