# round 2, call G: B half-step of the fast path at 5 / 6 CTAs per SM (lean kernel)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1; tail -n 2 gpurun_out/g_build.log
timeout 600 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal" > gpurun_out/g_pytest.log 2>&1
tail -n 3 gpurun_out/g_pytest.log
launches() { name=$1; shift
  env $ENVV timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/g_launches_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/g_ncu_launch.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/g_launches_$name.csv')) if len(r)>10 and r[0].isdigit()]
pl=[(r[4][5:40],float(r[-1])/1e3) for r in rows if 'step3_plain' in r[4]]
print('$name plain launches (us), last 8:', [(k,round(x,1)) for k,x in pl[-4:]])
PY
}
ENVV="MEEP_B200_NOEPI_OCC=5" launches occ5
ENVV="MEEP_B200_NOEPI_OCC=6" launches occ6
ENVV="MEEP_B200_NOEPI_OCC=5" MEEP_B200_TEST=1 timeout 600 python -m pytest tests -m gpu -q -x -k "c2_3d_pml or 3d_metal" > gpurun_out/g_pytest5.log 2>&1
tail -n 2 gpurun_out/g_pytest5.log
ENVV="MEEP_B200_NOEPI_OCC=5" launches occ5_f32 --prec f32
ENVV="MEEP_B200_NOEPI_OCC=6" launches occ6_f32 --prec f32
ENVV="X=1" launches occ0_f32 --prec f32
