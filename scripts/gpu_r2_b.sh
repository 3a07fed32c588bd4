# round 2, call B: host-side rework (lazy host arrays, analytic connect, no double upload) on the device:
# parity subset, first-step timeline at 512^3, and the first 1024^3 run on ONE GPU
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1; tail -n 2 gpurun_out/b_build.log
timeout 900 python -m pytest tests -m gpu -q -x -k "c2_3d_pml or 3d_metal or bloch or lorentz_3d or midrun or phase_in or c3_au or cyl_m1 or sync_magnetic or known_results or three_d" > gpurun_out/b_pytest.log 2>&1
tail -n 5 gpurun_out/b_pytest.log
MEEP_B200_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_512.json 2> gpurun_out/b_bench_512.err
grep -v "no E/H fusion\|recorded\|phase " gpurun_out/b_bench_512.err | cut -c1-200 | tail -n 30
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench_512.json').read().strip().splitlines()[-1])
print('512:', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2), d['probe']['values'][:3])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
MEEP_B200_VERBOSE=1 MEEP_B200_BENCH_N1=1024 /usr/bin/time -v timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_1024.json 2> gpurun_out/b_bench_1024.err
grep -v "no E/H fusion\|recorded\|phase " gpurun_out/b_bench_1024.err | cut -c1-200 | tail -n 45
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench_1024.json').read().strip().splitlines()[-1])
print('1024:', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2), 'frac', d['roofline']['whole_step']['frac'], d['probe'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
PY
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_c3.json 2> gpurun_out/b_bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench_c3.json').read().strip().splitlines()[-1])
print('c3:', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s', 'frac', d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['whole_step'])
print({k:(round(v['ms_per_step'],3), round(v['alg_bytes_per_step']/1e9,3)) for k,v in d['roofline']['kernels'].items()})
PY
