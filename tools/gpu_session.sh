#!/bin/bash
# scratch driver for one gpurun call: ./tools/gpu_session.sh <stage ...>; logs under gpurun_out/
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print("%7.1f Mpix/s step %.3f ms filter %.3f ms prepass %.3f ms  %s  parity=%s  e2e=%s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_prepass"]["kernel_ms"], d["config"]["kernel"], json.dumps(d.get("parity") and {k: d["parity"][k] for k in ("rel_mad","flips","ok","timed_plan_bit_identical")}), d.get("e2e") and round(d["e2e"]["value"],1)))'
QB="python bench.py --no-cpu-baseline --no-accum --no-8k --no-acrr"
for stage in "$@"; do
  echo "=== $stage"
  case $stage in
    tests_denoiser) timeout 900 python -m pytest tests/test_denoiser_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -15 ;;
    tests_all) timeout 2400 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 ;;
    bench_sym) timeout 600 $QB --no-e2e --steps 10 2>gpurun_out/bench_sym.err | tee gpurun_out/bench_sym.json | python -c "$P" || tail -5 gpurun_out/bench_sym.err ;;
    bench_stream) SMC_FILTER_KERNEL=stream timeout 600 $QB --no-e2e --steps 10 2>gpurun_out/bench_stream.err | tee gpurun_out/bench_stream.json | python -c "$P" || tail -5 gpurun_out/bench_stream.err ;;
    bench_e2e) timeout 600 $QB --steps 10 2>gpurun_out/bench_e2e.err | tee gpurun_out/bench_e2e.json | python -c "$P" || tail -5 gpurun_out/bench_e2e.err ;;
    bench_full) timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_full.err | tee gpurun_out/bench_full.json | python -c "$P" || tail -5 gpurun_out/bench_full.err ;;
    bench_ref) timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-900 ;;
    bench_acrr) timeout 600 python bench.py --no-cpu-baseline --no-accum --no-8k --no-e2e --no-parity --steps 5 2>gpurun_out/bench_acrr.err | tee gpurun_out/bench_acrr.json | python -c 'import sys,json; print(json.dumps(json.loads(sys.stdin.readlines()[-1])["acrr"]))' || tail -5 gpurun_out/bench_acrr.err ;;
    driver) # what the driver runs at round end: smoke, the reference arm, the default bench line
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
      timeout 900 python bench.py --impl reference 2>gpurun_out/drv_ref.err | tee gpurun_out/drv_ref.json | cut -c1-300
      timeout 900 python bench.py 2>gpurun_out/drv_bench.err | tee gpurun_out/drv_bench.json | python -c "$P" || tail -5 gpurun_out/drv_bench.err ;;
    sweep_*) # sweep_<ENVVAR>=v1:v2:...   device-resident bench per value
      kv=${stage#sweep_}; var=${kv%%=*}; vals=${kv#*=}
      for v in ${vals//:/ }; do echo -n "[$var=$v] "; env $var=$v timeout 300 $QB --no-e2e --no-parity --steps 10 2>gpurun_out/sweep.err | python -c "$P" || tail -3 gpurun_out/sweep.err; done ;;
    accum) timeout 900 python tools/bench_accum.py --json gpurun_out/bench_accum.json 4 16 64 256 2048 2>&1 | tail -20 ;;
    accum_quick) timeout 600 python tools/bench_accum.py --dist heavy 16 64 256 2>&1 | tail -8 ;;
    tests_moments) timeout 900 python -m pytest tests/test_moments_gpu.py tests/test_configs_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -8 ;;
    render) timeout 1500 python tools/bench_ref_render.py --out gpurun_out/ref_render.json 2>&1 | tail -3 | cut -c1-1500 ;;
    probe) python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 tools/probe_h2d.py --out gpurun_out/probe_h2d_n${NGPU:-2}.json 2>&1 | tail -40 ;;
    mgpu_tests) timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q --timeout 900 2>&1 | tail -8 ;;
    mgpu_bench) for n in ${NLIST:-2}; do echo "--- N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/bench_n$n.err | tee gpurun_out/bench_n$n.json | python -c "$P" || tail -5 gpurun_out/bench_n$n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
k=d.get("config4_8k") or {}
print("   8k: %s Mpix/s %s ms parity=%s" % (k.get("value"), k.get("ms_per_step"), (k.get("parity") or {}).get("ok")))
print("   e2e: %s" % (d.get("e2e") and d["e2e"]["value"]))
PY
done ;;
    pipe_trace) SMC_PIPE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${NGPU:-2} --steps 3 --warmup 3 --no-8k --no-accum 2> gpurun_out/pipe_trace_n${NGPU:-2}.txt | tail -1 | cut -c1-300 ;;
    ncu_step) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"filter_sym|sym_gather|prepass_kernel|nonfinite_fixup" -s 12 -c 4 -f -o gpurun_out/prof_step $QB --no-e2e --no-parity --steps 1 --warmup 3 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log ;;
    ncu_accum) timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_stream -s 2 -c 1 -f -o gpurun_out/prof_accum python tools/bench_accum.py --dist heavy 16 > gpurun_out/ncu_accum.log 2>&1; tail -2 gpurun_out/ncu_accum.log ;;
    ncu_accum64) timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_stream -s 2 -c 1 -f -o gpurun_out/prof_accum64 python tools/bench_accum.py --dist heavy 64 > gpurun_out/ncu_accum64.log 2>&1; tail -2 gpurun_out/ncu_accum64.log ;;
    ref_estimator) timeout 600 python tools/bench_ref_estimator.py 2>&1 | tail -1 | tee gpurun_out/ref_estimator_4k.json ;;
    ref_estimator_ab) for v in 0 1 0 1; do echo -n "[defer=$v] "; STATMC_B200_DEFER_UPLOADS=$v timeout 600 python tools/bench_ref_estimator.py 2>&1 | tail -1 | tee gpurun_out/ref_estimator_4k_defer$v.json | cut -c100-260; done ;;
    sanitizer) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_denoiser_gpu.py -m gpu -x -q --timeout 1200 -k "rgb_default or large_radius or gbuffer_sets or tiny or peer_halo or host_pipelined or nonfinite or triples or device_table" 2>&1 | tail -6 ;;
    ncu_filter) timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_sym -s 3 -c 1 -f -o gpurun_out/prof_filter $QB --no-e2e --no-parity --steps 1 --warmup 3 > gpurun_out/ncu_filter.log 2>&1; tail -2 gpurun_out/ncu_filter.log ;;
    ncu_launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $QB --no-e2e --no-parity --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
