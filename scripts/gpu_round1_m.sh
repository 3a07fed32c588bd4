set -x
mkdir -p gpurun_out
free -g > gpurun_out/m_mem.txt; nproc >> gpurun_out/m_mem.txt
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
MEEP_B200_PEER_TIMEOUT_S=120 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --size 1024 --scaling strong --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_2gpu_strong_1024.json 2> gpurun_out/m_bench_2gpu_strong_1024.err
cat gpurun_out/m_bench_2gpu_strong_1024.json
tail -n 3 gpurun_out/m_bench_2gpu_strong_1024.err
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail" -gt 300 ]; then
  timeout 1500 python bench.py --size 1024 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_1gpu_1024.json 2> gpurun_out/m_bench_1gpu_1024.err
  cat gpurun_out/m_bench_1gpu_1024.json
  tail -n 5 gpurun_out/m_bench_1gpu_1024.err
else
  echo "only $avail GB of host RAM: no single-GPU 1024^3 run" | tee gpurun_out/m_bench_1gpu_1024.err
fi
