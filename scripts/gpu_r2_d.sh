# round 2, call D: fast-path kernel at 4 CTAs/SM (64 registers): A/B of the per-job form, the
# two-planes-in-flight B half, single precision; per-launch times of both halves; first-step trace
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1; tail -n 2 gpurun_out/d_build.log
timeout 600 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal" > gpurun_out/d_pytest.log 2>&1
tail -n 3 gpurun_out/d_pytest.log
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/d_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'warm', round(d['config']['warmup_s'],1))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/d_bench_$name.err').read()[-1500:])
PY
}
ENVV="MEEP_B200_VERBOSE=1" run 512_default
grep "took\|context\|scan\|upload" gpurun_out/d_bench_512_default.err | head -40
ENVV="MEEP_B200_PAIR_PLANES=0" run 512_nopair
ENVV="MEEP_B200_PLAIN_PER_JOB=0" run 512_tablejob
ENVV="X=1" run 512_f32 --prec f32
ENVV="MEEP_B200_PAIR_PLANES=0" run 512_f32_nopair --prec f32
for v in 1 0; do
MEEP_B200_PAIR_PLANES=$v timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d_launches_pair$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu_launch.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/d_launches_pair$v.csv')) if len(r)>10 and r[0].isdigit()]
pl=[float(r[-1])/1e3 for r in rows if 'step3_plain' in r[4]]
print('pair=$v plain launches (us), last 8:', [round(x,1) for x in pl[-8:]])
PY
done
