# round 2, call X (1 GPU): off-diagonal E update with the three component jobs interleaved tile by tile (A/B) + parity
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/x_build.log 2>&1; tail -n 2 gpurun_out/x_build.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/x_bench_$name.json 2> gpurun_out/x_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/x_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/x_bench_%s.err'%n).read()[-1500:])
PY
}
run c4 -- --workload c4
run c4_seq MEEP_B200_EDHB_INTERLEAVE=0 -- --workload c4
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/x_pytest_kernels.log 2>&1; tail -n 3 gpurun_out/x_pytest_kernels.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "test_b200_matches_reference and (aniso or offdiag or chi3 or c4)" > gpurun_out/x_pytest_parity.log 2>&1; tail -n 3 gpurun_out/x_pytest_parity.log
