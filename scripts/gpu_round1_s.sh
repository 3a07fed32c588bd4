set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; echo "smoke rc=$?"
tail -n 2 gpurun_out/s_smoke.log
timeout 900 python bench.py > gpurun_out/s_bench_default.json 2> gpurun_out/s_bench_default.err; echo "bench rc=$?"
cat gpurun_out/s_bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/s_bench_reference.json 2> gpurun_out/s_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/s_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s_launches_c2_512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s_ncu_launch.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/s_pytest.log
