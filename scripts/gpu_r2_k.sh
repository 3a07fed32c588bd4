# round 2, call K (8 GPUs): the driver's own multi-GPU commands — 1024^3 strong scaling on 8 and 4 B200
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/k_topo.txt 2>&1
nproc; free -g | head -2
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/k_build.log 2>&1; tail -n 2 gpurun_out/k_build.log
show() { python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/k_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', d['n_gpus'], 'gpus', d['scaling'], d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    print('    probe', d['probe']['after_steps'], d['probe']['values'])
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/k_bench_$1.err').read()[-3000:])
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/k_bench_8gpu.json 2> gpurun_out/k_bench_8gpu.err
show 8gpu
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/k_bench_4gpu.json 2> gpurun_out/k_bench_4gpu.err
show 4gpu
