# round 2, call E: what HBM delivers per read:write mix (stream_mix), fast-path kernel variants per launch
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1; tail -n 2 gpurun_out/e_build.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a bench/micro/stream_mix.cu -o gpurun_out/stream_mix && ./gpurun_out/stream_mix > gpurun_out/e_stream_mix.jsonl; cat gpurun_out/e_stream_mix.jsonl
timeout 600 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal or ld_preload" > gpurun_out/e_pytest.log 2>&1
tail -n 3 gpurun_out/e_pytest.log
launches() { name=$1
  env $ENVV timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/e_launches_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu_launch.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/e_launches_$name.csv')) if len(r)>10 and r[0].isdigit()]
pl=[float(r[-1])/1e3 for r in rows if 'step3_plain' in r[4]]
print('$name plain launches (us), last 8:', [round(x,1) for x in pl[-8:]])
PY
}
ENVV="X=1" launches default
ENVV="MEEP_B200_PAIR_PLANES=0" launches nopair
ENVV="MEEP_B200_PLAIN_FAST=0" launches oldloop_perjob
ENVV="MEEP_B200_PLAIN_FAST=0 MEEP_B200_PLAIN_PER_JOB=0" launches oldloop_table
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/e_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'warm', round(d['config']['warmup_s'],1))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/e_bench_$name.err').read()[-1500:])
PY
}
ENVV="X=1" run 512_default
ENVV="X=1" run 512_f32 --prec f32
ENVV="MEEP_B200_PLAIN_FAST=0" run 512_f32_oldloop --prec f32
