set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/d_gpus.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --maxfail=5 > gpurun_out/d_pytest_mp.log 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --n 256 --steps 20 --warmup 3 > gpurun_out/d_bench_2gpu_256.json 2> gpurun_out/d_bench_2gpu_256.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/d_bench_2gpu_512.json 2> gpurun_out/d_bench_2gpu_512.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/d_bench_2gpu_512_strong.json 2> gpurun_out/d_bench_2gpu_512_strong.err
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench_512.json 2> gpurun_out/d_bench_512.err
tail -n 5 gpurun_out/d_pytest_mp.log
tail -n 3 gpurun_out/d_bench_2gpu_256.err
