set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/h_bench_2gpu_weak_512.json 2> gpurun_out/h_bench_2gpu_weak_512.err
timeout 900 python bench.py --workload c3 --size 320 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_c3_320.json 2> gpurun_out/h_bench_c3_320.err
timeout 900 python -m pytest tests -m gpu -q -k "lorentz or au_sphere or polariton or multigpu or beta or sym" > gpurun_out/h_pytest.log 2>&1
tail -n 3 gpurun_out/h_pytest.log
