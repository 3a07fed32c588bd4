set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/e_bench_2gpu_512.json 2> gpurun_out/e_bench_2gpu_512.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/e_bench_2gpu_512_strong.json 2> gpurun_out/e_bench_2gpu_512_strong.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 20 --warmup 3 --size 256 > gpurun_out/e_bench_2gpu_256.json 2> gpurun_out/e_bench_2gpu_256.err
tail -n 3 gpurun_out/e_bench_2gpu_512.err
