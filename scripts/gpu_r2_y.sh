# round 2, call Y (1 GPU): PML kernel budget / CTAs-per-SM A/B in the product (512^3, double and single)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/y_build.log 2>&1; tail -n 2 gpurun_out/y_build.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/y_bench_$name.json 2> gpurun_out/y_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/y_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/y_bench_%s.err'%n).read()[-1500:])
PY
}
for v in 31 32 41 42; do run f64_$v MEEP_B200_SPLIT_PML=$v -- --n 512; done
for v in 4 31 32 41 42; do run f32_$v MEEP_B200_SPLIT_PML=$v -- --n 512 --prec f32; done
