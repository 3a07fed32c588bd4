# round 2, final call 3 (4 GPUs): the driver's 8-GPU command on the final build
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f4_build.log 2>&1; tail -n 2 gpurun_out/f4_build.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/f4_bench_4gpu.json 2> gpurun_out/f4_bench_4gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/f4_bench_4gpu.json').read().strip().splitlines()[-1])
    print('4gpu', d['scaling'], d['config']['cell'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'e2e', round(d['e2e']['value']/1e9,2))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
    print('    probe', d['probe']['values'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/f4_bench_4gpu.err').read()[-2500:])
PY
