set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/x_build.log 2>&1
run() { name=$1; shift
  env $ENVV timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/x_bench_$name.json 2> gpurun_out/x_bench_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/x_bench_$name.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'],3), round(d['value']/1e9,2), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
}
ENVV="MEEP_B200_PLAIN_UNROLL=2" run f32_u2 --prec f32
ENVV="MEEP_B200_PLAIN_UNROLL=24" run f32_u2_occ4 --prec f32
ENVV="MEEP_B200_PLAIN_UNROLL=2" run f64_u2
