# round 2, call I (1 GPU): full GPU test suite, the driver's own N=1 command, ncu captures for the roofline
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1; tail -n 2 gpurun_out/i_build.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i_pytest.log 2>&1
tail -n 6 gpurun_out/i_pytest.log
MEEP_B200_VERBOSE=1 timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
grep "took\|context\|scan\|upload\|find_metals\|bench:" gpurun_out/i_bench_default.err | head -30
python - <<'PY'
import json
d=json.loads(open('gpurun_out/i_bench_default.json').read().strip().splitlines()[-1])
print('default', d['n_gpus'], d['scaling'], d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'rss', round(d['config']['host_max_rss_gb'],1), 'e2e', round(d['e2e']['value']/1e9,2))
print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
print('    probe', d['probe']['after_steps'], d['probe']['values'])
print('    cpu', d.get('cpu_baseline'))
c=d.get('configs1_512'); print('    512:', c and (round(c['ms_per_step'],3), round(c['value']/1e9,2), c['roofline']['whole_step']['frac']))
PY
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 | cut -c1-600
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:step3_plain -s 10 -c 2 -o gpurun_out/i_prof_plain_1024 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_plain.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:dft_kernel -s 4 -c 2 -o gpurun_out/i_prof_dft_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_dft.log 2>&1
ls -la gpurun_out/i_*
