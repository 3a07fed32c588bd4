set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_512_$name.json 2> gpurun_out/q_bench_512_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/q_bench_512_$name.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
}
run occ4_t1_8 MEEP_B200_SPLIT_PML=4 MEEP_B200_PML_T1=8
run occ4_t1_16 MEEP_B200_SPLIT_PML=4 MEEP_B200_PML_T1=16
run occ4_t1_32 MEEP_B200_SPLIT_PML=4 MEEP_B200_PML_T1=32
run occ5_t1_8 MEEP_B200_SPLIT_PML=5 MEEP_B200_PML_T1=8
run occ5_t1_16 MEEP_B200_SPLIT_PML=5 MEEP_B200_PML_T1=16
run occ6_t1_16 MEEP_B200_SPLIT_PML=6 MEEP_B200_PML_T1=16
