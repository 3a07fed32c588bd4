# round 2, final call 2 (2 GPUs): sharded parity on hardware with the final code, the driver's 2-GPU command
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f2_build.log 2>&1; tail -n 2 gpurun_out/f2_build.log
nvidia-smi topo -m > gpurun_out/f2_topo_2gpu.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --durations=5 > gpurun_out/f2_pytest_mgpu.log 2>&1
tail -n 12 gpurun_out/f2_pytest_mgpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/f2_bench_2gpu.json 2> gpurun_out/f2_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/f2_bench_2gpu.json').read().strip().splitlines()[-1])
    print('2gpu', d['scaling'], d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    print('    probe', d['probe']['values'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/f2_bench_2gpu.err').read()[-2500:])
PY
