set -x
mkdir -p gpurun_out
free -g > gpurun_out/j_mem.txt; nproc >> gpurun_out/j_mem.txt
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -k "cyl or golden or kernels or reference_test_program" > gpurun_out/j_pytest_cyl.log 2>&1
tail -n 5 gpurun_out/j_pytest_cyl.log
# 1024^3: the north-star target size on one GPU (double precision, ~93 GB resident); needs ~2x that in host RAM
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail" -gt 330 ]; then
  timeout 1500 python bench.py --size 1024 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_1024.json 2> gpurun_out/j_bench_1024.err
  cat gpurun_out/j_bench_1024.json
  tail -n 5 gpurun_out/j_bench_1024.err
else
  echo "only $avail GB of host RAM available: skipping 1024^3 on one GPU" | tee gpurun_out/j_bench_1024.err
  timeout 1200 python bench.py --size 768 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_768.json 2> gpurun_out/j_bench_768.err
  cat gpurun_out/j_bench_768.json
fi
