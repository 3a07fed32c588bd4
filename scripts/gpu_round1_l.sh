set -x
mkdir -p gpurun_out
N=${1:-8}
free -g > gpurun_out/l_mem_${N}.txt; nproc >> gpurun_out/l_mem_${N}.txt
nvidia-smi topo -m > gpurun_out/l_topo_${N}.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/l_build.log 2>&1
# the north-star target case: 1024^3 dielectric + PML, strong scaling over N GPUs
MEEP_B200_PEER_TIMEOUT_S=120 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --size 1024 --scaling strong --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench_${N}gpu_strong_1024.json 2> gpurun_out/l_bench_${N}gpu_strong_1024.err
cat gpurun_out/l_bench_${N}gpu_strong_1024.json
tail -n 3 gpurun_out/l_bench_${N}gpu_strong_1024.err
