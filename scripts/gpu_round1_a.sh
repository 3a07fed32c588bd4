set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 600 python -c "import __graft_entry__ as g; g.build(); print('BUILD_OK')" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -x -k "kernels" > gpurun_out/pytest_kernels.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 -k "not kernels" > gpurun_out/pytest_parity.log 2>&1
timeout 900 python bench.py --n 256 --steps 20 --warmup 3 --cpu-n 128 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_256.csv python bench.py --n 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3 -s 6 -c 2 -o gpurun_out/prof_step3 python bench.py --n 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_kernels.log gpurun_out/smoke.log gpurun_out/pytest_parity.log; cat gpurun_out/bench_256.json gpurun_out/bench_512.json | cut -c1-600
