# round 2, call H (2 GPUs): sharded parity on hardware (peer-memory and NCCL transports), the driver's
# own multi-GPU command (default: 1024^3 strong), 512^3 strong
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1; tail -n 2 gpurun_out/h_build.log
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/h_pytest_mgpu.log 2>&1
tail -n 4 gpurun_out/h_pytest_mgpu.log
show() { python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/h_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', d['n_gpus'], 'gpus', d['scaling'], d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    print('    probe', d['probe']['after_steps'], d['probe']['values'][:4])
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/h_bench_$1.err').read()[-2500:])
PY
}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/h_bench_2gpu_default.json 2> gpurun_out/h_bench_2gpu_default.err
show 2gpu_default
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --size 512 > gpurun_out/h_bench_2gpu_strong_512.json 2> gpurun_out/h_bench_2gpu_strong_512.err
show 2gpu_strong_512
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_1gpu_512.json 2> gpurun_out/h_bench_1gpu_512.err
show 1gpu_512
