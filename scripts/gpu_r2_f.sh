# round 2, call F: cp.async-staged fast path: parity, per-launch times, bench (f64, f32), first-step trace
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1; tail -n 2 gpurun_out/f_build.log
timeout 600 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal or 3d_bloch or sync_magnetic" > gpurun_out/f_pytest.log 2>&1
tail -n 3 gpurun_out/f_pytest.log
launches() { name=$1; shift
  env $ENVV timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/f_launches_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/f_ncu_launch.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/f_launches_$name.csv')) if len(r)>10 and r[0].isdigit()]
pl=[float(r[-1])/1e3 for r in rows if 'step3_plain' in r[4]]
print('$name plain launches (us), last 8:', [round(x,1) for x in pl[-8:]])
PY
}
ENVV="X=1" launches piped
ENVV="MEEP_B200_PLAIN_PIPED=0 MEEP_B200_PLAIN_FAST=0" launches oldloop_perjob
ENVV="X=1" launches piped_f32 --prec f32
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/f_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'warm', round(d['config']['warmup_s'],1))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/f_bench_$name.err').read()[-1500:])
PY
}
ENVV="MEEP_B200_VERBOSE=1" run 512_piped
grep "took\|context\|scan\|upload\|find_metals" gpurun_out/f_bench_512_piped.err | head -40
ENVV="X=1" run 512_f32_piped --prec f32
ENVV="MEEP_B200_PLAIN_PIPED=0 MEEP_B200_PLAIN_FAST=0 MEEP_B200_PLAIN_PER_JOB=0" run 512_round1_form
