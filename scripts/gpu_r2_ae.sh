# round 2, call AE (1 GPU): lean + shell fast path as the default: parity, DRAM traffic of one step's fast-path
# launches (metrics-only ncu pass: dram bytes, duration, registers, grid), the driver's default command, single precision
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ae_build.log 2>&1; tail -n 2 gpurun_out/ae_build.log
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "test_b200_matches_reference or opt_in or unfused or kernels" > gpurun_out/ae_pytest_parity.log 2>&1; tail -n 3 gpurun_out/ae_pytest_parity.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:"step3_lean|step3_plain|step3_cols" -s 16 -c 9 -o /tmp/ae_fast_1024 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/ae_ncu_fast.log 2>&1
cp /tmp/ae_fast_1024.ncu-rep gpurun_out/ae_prof_fast_path_1024.ncu-rep
python scripts/ncu_traffic.py gpurun_out/ae_prof_fast_path_1024.ncu-rep c2 1024 f64 step3 && cp profiles/traffic_c2_1024_f64.json gpurun_out/ae_traffic_c2_1024_f64.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/ae_bench_default.json 2> gpurun_out/ae_bench_default.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --n 512 --prec f32 > gpurun_out/ae_bench_512_f32.json 2> gpurun_out/ae_bench_512_f32.err
python - <<'PY'
import json
for n in ['default','512_f32']:
    try:
        d=json.loads(open('gpurun_out/ae_bench_%s.json'%n).read().strip().splitlines()[-1])
        r=d.get('roofline',{})
        print(n, round(d.get('ms_per_step',0),3), 'ms', round(d['value']/1e9,3), 'Gc/s', 'whole', round(r.get('whole_step',{}).get('frac',0),3), 'dom', round(r.get('frac',0),3), 'traffic', r.get('traffic'), 'alg', r.get('alg_bytes_per_launch'), 'traffic_ok', r.get('traffic_build_is_this_build'), {k:round(v['ms_per_step'],3) for k,v in r.get('kernels',{}).items()}, 'launches', d.get('gpu_launches'))
        if 'configs1_512' in d: print('   configs1_512', round(d['configs1_512']['ms_per_step'],3), round(d['configs1_512']['roofline']['whole_step']['frac'],3))
        if 'probe' in d: print('   probe', d['probe']['component'], d['probe']['values'][:4])
    except Exception as e:
        print(n, 'FAILED', e)
PY
