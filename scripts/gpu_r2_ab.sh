# round 2, call AB (1 GPU): fast path as lean + rest launches: harness table, parity, product A/B (512^3 double / single)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ab_build.log 2>&1; tail -n 2 gpurun_out/ab_build.log
timeout 200 bench/micro/_build/pml_shapes 492 10 lean > gpurun_out/ab_lean.jsonl 2> gpurun_out/ab_lean.err; cat gpurun_out/ab_lean.err
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/ab_pytest_kernels.log 2>&1; tail -n 3 gpurun_out/ab_pytest_kernels.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "test_b200_matches_reference or chunk_count" > gpurun_out/ab_pytest_parity.log 2>&1; tail -n 3 gpurun_out/ab_pytest_parity.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'probe', d['probe']['values'][:2])
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/ab_bench_%s.err'%n).read()[-1500:])
PY
}
run f64_lean -- --n 512
run f64_masked MEEP_B200_PLAIN_LEAN=0 -- --n 512
run f32_lean -- --n 512 --prec f32
run f32_masked MEEP_B200_PLAIN_LEAN=0 -- --n 512 --prec f32
