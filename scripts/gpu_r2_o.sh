# round 2, call O (1 GPU): full GPU suite on the final kernels, c3 / c4 / f32 lines, the driver's N=1 command,
# traffic capture of the dominant kernel on this build
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/o_build.log 2>&1; tail -n 2 gpurun_out/o_build.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/o_pytest.log 2>&1
tail -n 5 gpurun_out/o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/o_bench_$name.json 2> gpurun_out/o_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/o_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_build_is_this_build'), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    c=d.get('configs1_512'); print('    512:', c and (round(c['ms_per_step'],3), round(c['value']/1e9,2), round(c['roofline']['whole_step']['frac'],3)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/o_bench_$name.err').read()[-1500:])
PY
}
ENVV="X=1" run c3 --workload c3 --steps 40 --no-cpu-baseline
ENVV="X=1" run c4 --workload c4 --no-cpu-baseline
ENVV="X=1" run 512_f32 --size 512 --prec f32 --no-cpu-baseline
ENVV="X=1" run default
timeout 1200 ncu --set full --clock-control none -k regex:step3_plain -s 10 -c 2 -o /tmp/o_plain_1024 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/o_ncu_plain.log 2>&1
cp /tmp/o_plain_1024.ncu-rep gpurun_out/o_prof_plain_1024.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/o_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/o_ncu_launch.log 2>&1
du -sh gpurun_out
