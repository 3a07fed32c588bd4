set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/t_build.log 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/t_pytest_mgpu.log 2>&1; echo "mgpu rc=$?"
tail -n 3 gpurun_out/t_pytest_mgpu.log
timeout 900 python -m pytest tests -m gpu -q -x -k "gyro or aniso_sigma or halo_zero or golden or sync_magnetic" > gpurun_out/t_pytest_new.log 2>&1; echo "new rc=$?"
tail -n 3 gpurun_out/t_pytest_new.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/t_bench_2gpu_weak_512.json 2> gpurun_out/t_bench_2gpu_weak_512.err; echo "bench rc=$?"
cat gpurun_out/t_bench_2gpu_weak_512.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/t_bench_2gpu_reference.json 2> gpurun_out/t_bench_2gpu_reference.err; echo "ref rc=$?"
cat gpurun_out/t_bench_2gpu_reference.json
