# round 2, call AD (1 GPU): fast path as lean + shell launches (full box / x-slabs / shell columns): harness, parity, product lines
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ad_build.log 2>&1; tail -n 2 gpurun_out/ad_build.log
timeout 100 bench/micro/_build/pml_shapes 492 10 lean 2> gpurun_out/ad_lean.err | grep -E "masked|lean_plus_shell" > gpurun_out/ad_lean.jsonl; cat gpurun_out/ad_lean.err gpurun_out/ad_lean.jsonl
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean_plus_rest" > gpurun_out/ad_pytest_lean.log 2>&1; tail -n 3 gpurun_out/ad_pytest_lean.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 12 --warmup 4 "$@" > gpurun_out/ad_bench_$name.json 2> gpurun_out/ad_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ad_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'probe', d['probe']['values'][:2])
    if 'configs1_512' in d: print('   configs1_512', round(d['configs1_512']['ms_per_step'],3))
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/ad_bench_%s.err'%n).read()[-1500:])
PY
}
run f32_lean MEEP_B200_PLAIN_LEAN=1 -- --n 512 --prec f32
run f64_1024_lean MEEP_B200_PLAIN_LEAN=1 --
