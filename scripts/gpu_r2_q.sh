# round 2, call Q (1 GPU): durations of the GPU suite, traffic capture of the dominant kernel on the final build
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1; tail -n 2 gpurun_out/q_build.log
timeout 1200 ncu --set full --clock-control none -k regex:step3_plain -s 10 -c 2 -o /tmp/q_plain_1024 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_plain.log 2>&1
cp /tmp/q_plain_1024.ncu-rep gpurun_out/q_prof_plain_1024.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q -x --durations=30 > gpurun_out/q_pytest.log 2>&1
tail -n 40 gpurun_out/q_pytest.log
