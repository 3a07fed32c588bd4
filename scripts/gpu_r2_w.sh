# round 2, call W (1 GPU): f_minus_p kernel A/B (two elements in flight at four CTAs per SM vs element by element)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/w_build.log 2>&1; tail -n 2 gpurun_out/w_build.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/w_bench_$name.json 2> gpurun_out/w_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/w_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/w_bench_%s.err'%n).read()[-1500:])
PY
}
run c3 -- --workload c3
run c3_simple MEEP_B200_FMP_SIMPLE=1 -- --workload c3
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/w_pytest_kernels.log 2>&1; tail -n 3 gpurun_out/w_pytest_kernels.log
