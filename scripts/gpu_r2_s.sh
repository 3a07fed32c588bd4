# round 2, call S (1 GPU): PML kernel with a register budget spent on planes in flight (2/3/4 CTAs per SM),
# ncu of single chunk shapes (thin face vs thick cube), bench lines at 3 and 2 CTAs per SM
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1; tail -n 2 gpurun_out/s_build.log
timeout 600 bench/micro/_build/pml_shapes > gpurun_out/s_pml_shapes.jsonl 2> gpurun_out/s_pml_shapes.err; tail -n 3 gpurun_out/s_pml_shapes.err
for sh in face_x face_z cube_x all26; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:step3c -s 3 -c 1 -o /tmp/s_prof_$sh bench/micro/_build/pml_shapes 492 10 $sh c3_b20 16 > gpurun_out/s_ncu_$sh.log 2>&1
  cp /tmp/s_prof_$sh.ncu-rep gpurun_out/s_prof_pml_$sh.ncu-rep
done
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/s_bench_$name.json 2> gpurun_out/s_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/s_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/s_bench_%s.err'%n).read()[-1500:])
PY
}
run f64_c4 -- --n 512
run f64_c3 MEEP_B200_SPLIT_PML=3 -- --n 512
run f64_c2 MEEP_B200_SPLIT_PML=2 -- --n 512
run f32_c4 -- --n 512 --prec f32
run f32_c3 MEEP_B200_SPLIT_PML=3 -- --n 512 --prec f32
run f32_c2 MEEP_B200_SPLIT_PML=2 -- --n 512 --prec f32
