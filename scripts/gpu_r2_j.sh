# round 2, call J (1 GPU): PML kernel with two planes in flight (A/B), halo kernel with one run per warp
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1; tail -n 2 gpurun_out/j_build.log
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal or 3d_bloch or cond_chi3 or bfast or xperiodic or 3d_tiled or phase_in" > gpurun_out/j_pytest.log 2>&1
tail -n 3 gpurun_out/j_pytest.log
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/j_bench_$name.json 2> gpurun_out/j_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/j_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'warm', round(d['config']['warmup_s'],1))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/j_bench_$name.err').read()[-1500:])
PY
}
ENVV="X=1" run 512_pmlpair --size 512
ENVV="MEEP_B200_PML_PAIR=0" run 512_pmlsingle --size 512
ENVV="X=1" run 512_f32_pmlpair --size 512 --prec f32
ENVV="MEEP_B200_PML_PAIR=0" run 512_f32_pmlsingle --size 512 --prec f32
ENVV="X=1" run c3 --workload c3 --steps 40
ENVV="X=1" run c4 --workload c4
