set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/i_topo.txt 2>&1
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv >> gpurun_out/i_topo.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1
# multi-process parity first (peer-memory exchange on, then NCCL)
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/i_pytest_mgpu_p2p.log 2>&1
tail -n 3 gpurun_out/i_pytest_mgpu_p2p.log
MEEP_B200_P2P=0 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/i_pytest_mgpu_nccl.log 2>&1
tail -n 3 gpurun_out/i_pytest_mgpu_nccl.log
# 2-GPU weak, peer-memory exchange
MEEP_B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/i_bench_2gpu_weak_512_p2p.json 2> gpurun_out/i_bench_2gpu_weak_512_p2p.err
cat gpurun_out/i_bench_2gpu_weak_512_p2p.json
# same over NCCL with transport diagnostics
MEEP_B200_P2P=0 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/i_bench_2gpu_weak_512_nccl.json 2> gpurun_out/i_bench_2gpu_weak_512_nccl.err
cat gpurun_out/i_bench_2gpu_weak_512_nccl.json
# strong scaling 512^3 on 2 GPUs, peer-memory
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/i_bench_2gpu_strong_512_p2p.json 2> gpurun_out/i_bench_2gpu_strong_512_p2p.err
cat gpurun_out/i_bench_2gpu_strong_512_p2p.json
# c3 with the 16-blocks-per-CTA Lorentz kernel
timeout 900 python bench.py --workload c3 --size 320 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_c3_320.json 2> gpurun_out/i_bench_c3_320.err
cat gpurun_out/i_bench_c3_320.json
timeout 600 python -m pytest tests -m gpu -q -k "lorentz or au_sphere or polariton" > gpurun_out/i_pytest_lorentz.log 2>&1
tail -n 3 gpurun_out/i_pytest_lorentz.log
