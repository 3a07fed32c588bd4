set -x
mkdir -p gpurun_out
N=${NG:-2}
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/g_bench_${N}gpu_weak_512.json 2> gpurun_out/g_bench_${N}gpu_weak_512.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 20 --warmup 3 --scaling strong > gpurun_out/g_bench_${N}gpu_strong_512.json 2> gpurun_out/g_bench_${N}gpu_strong_512.err
tail -n 3 gpurun_out/g_bench_${N}gpu_weak_512.err
