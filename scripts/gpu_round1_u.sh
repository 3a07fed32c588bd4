set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/u_build.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/u_bench_4gpu_weak_512.json 2> gpurun_out/u_bench_4gpu_weak_512.err; echo "bench rc=$?"
cat gpurun_out/u_bench_4gpu_weak_512.json
tail -n 3 gpurun_out/u_bench_4gpu_weak_512.err
