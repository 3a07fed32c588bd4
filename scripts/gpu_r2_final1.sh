# round 2, final call 1 (1 GPU): DRAM traffic of the dominant kernel on the final build, launch list, the driver's
# own commands (default line, reference arm), the other configurations, the full GPU suite
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f1_build.log 2>&1; tail -n 2 gpurun_out/f1_build.log
timeout 1200 ncu --set full --clock-control none -k regex:step3_plain -s 10 -c 2 -o /tmp/f1_plain_1024 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f1_ncu_plain.log 2>&1
cp /tmp/f1_plain_1024.ncu-rep gpurun_out/f1_prof_plain_1024.ncu-rep
python scripts/ncu_traffic.py gpurun_out/f1_prof_plain_1024.ncu-rep c2 1024 f64 step3 && cp profiles/traffic_c2_1024_f64.json gpurun_out/f1_traffic_c2_1024_f64.json
ncu -i /tmp/f1_plain_1024.ncu-rep --page details > gpurun_out/f1_step3_plain_1024_ncu_details.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches_default_1024.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f1_ncu_launch.log 2>&1
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f1_bench_default.json 2> gpurun_out/f1_bench_default.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/f1_bench_reference_arm.json 2> gpurun_out/f1_bench_reference_arm.err
for w in "512_f32 --n 512 --prec f32" "c3 --workload c3" "c4 --workload c4" "c3_f32 --workload c3 --prec f32"; do
  set -- $w; name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/f1_bench_$name.json 2> gpurun_out/f1_bench_$name.err
done
python - <<'PY'
import json
for n in ['default','reference_arm','512_f32','c3','c4','c3_f32']:
    try:
        d=json.loads(open('gpurun_out/f1_bench_%s.json'%n).read().strip().splitlines()[-1])
        r=d.get('roofline',{})
        print(n, round(d.get('ms_per_step',0),3), 'ms', round(d['value']/1e9,3), 'Gc/s', 'whole', round(r.get('whole_step',{}).get('frac',0),3), 'dom', round(r.get('frac',0),3), 'traffic_ok', r.get('traffic_build_is_this_build'), {k:round(v['ms_per_step'],3) for k,v in r.get('kernels',{}).items()})
        if 'configs1_512' in d: print('   configs1_512', round(d['configs1_512']['ms_per_step'],3), round(d['configs1_512']['roofline']['whole_step']['frac'],3))
        if 'probe' in d: print('   probe', d['probe']['component'], d['probe']['values'][:4])
    except Exception as e:
        print(n, 'FAILED', e)
PY
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/f1_pytest_gpu_full.log 2>&1
tail -n 20 gpurun_out/f1_pytest_gpu_full.log
