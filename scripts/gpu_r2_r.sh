# round 2, call R (1 GPU): PML kernel by chunk shape (bench/micro/pml_shapes), single-precision fast-path forms,
# two ranks sharing the one GPU (cross-process exchange), the slow reference test programs after the download fix
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r_build.log 2>&1; tail -n 2 gpurun_out/r_build.log
timeout 600 bench/micro/_build/pml_shapes > gpurun_out/r_pml_shapes.jsonl 2> gpurun_out/r_pml_shapes.err; tail -n 3 gpurun_out/r_pml_shapes.err
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/r_bench_$name.json 2> gpurun_out/r_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/r_bench_%s.err'%n).read()[-1500:])
PY
}
run f32_default -- --n 512 --prec f32
run f32_fast MEEP_B200_PLAIN_FAST=1 MEEP_B200_PLAIN_PER_JOB=1 -- --n 512 --prec f32
run f32_t1_32 MEEP_B200_PLAIN_T1=32 MEEP_B200_PML_T1=32 -- --n 512 --prec f32
run f64_default -- --n 512
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -k "sharing_one_gpu" --durations=5 > gpurun_out/r_pytest_onegpu.log 2>&1
tail -n 12 gpurun_out/r_pytest_onegpu.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "test_reference_test_program and flux" --durations=10 > gpurun_out/r_pytest_reftests.log 2>&1
tail -n 14 gpurun_out/r_pytest_reftests.log
