set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p_build.log 2>&1
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_512_$name.json 2> gpurun_out/p_bench_512_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/p_bench_512_$name.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
}
run default X=1
run t1_8 MEEP_B200_PML_T1=8
run t1_2 MEEP_B200_PML_T1=2
run occ4 MEEP_B200_SPLIT_PML=4
run occ4_t1_8 MEEP_B200_SPLIT_PML=4 MEEP_B200_PML_T1=8
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "sync_magnetic or c2_3d or bloch or metal" > gpurun_out/p_pytest.log 2>&1
tail -n 3 gpurun_out/p_pytest.log
