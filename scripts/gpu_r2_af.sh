# round 2, call AF (1 GPU): final-final build: DRAM traffic of one step's fast-path launches (metrics-only ncu pass),
# the driver's default command
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/af_build.log 2>&1; tail -n 2 gpurun_out/af_build.log
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:"step3_lean|step3_plain|step3_cols" -s 16 -c 9 -o /tmp/af_fast_1024 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/af_ncu_fast.log 2>&1
cp /tmp/af_fast_1024.ncu-rep gpurun_out/af_prof_fast_path_1024.ncu-rep
python scripts/ncu_traffic.py gpurun_out/af_prof_fast_path_1024.ncu-rep c2 1024 f64 step3 && cp profiles/traffic_c2_1024_f64.json gpurun_out/af_traffic_c2_1024_f64.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/af_bench_default.json 2> gpurun_out/af_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/af_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print('default', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,3), 'Gc/s whole', round(r['whole_step']['frac'],3), 'dom', round(r['frac'],3), 'traffic', r.get('traffic'), 'alg', r.get('alg_bytes_per_launch'), 'ok', r.get('traffic_build_is_this_build'), 'launches', d.get('gpu_launches'), 'e2e', round(d['e2e']['value']/1e9,3))
print('   configs1_512', round(d['configs1_512']['ms_per_step'],3), round(d['configs1_512']['roofline']['whole_step']['frac'],3), 'probe', d['probe']['values'][:3])
PY
