set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r_build.log 2>&1
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r_bench_$name.json 2> gpurun_out/r_bench_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r_bench_$name.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'],3), round(d['value']/1e9,2), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
}
ENVV="X=1" run 512_default
ENVV="MEEP_B200_PLAIN_T1=16" run 512_plain_t1_16
ENVV="MEEP_B200_PLAIN_T1=32" run 512_plain_t1_32
ENVV="MEEP_B200_PLAIN_T1=4" run 512_plain_t1_4
ENVV="X=1" run c4_512 --workload c4 --size 512
ENVV="X=1" run c3_320 --workload c3 --size 320
