set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "not multigpu" > gpurun_out/f_pytest.log 2>&1
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench_512.json 2> gpurun_out/f_bench_512.err
MEEP_B200_PARAMJOBS=1 timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_512_paramjobs.json 2> gpurun_out/f_bench_512_paramjobs.err
timeout 1500 python bench.py --workload c3 --size 320 --steps 20 --warmup 3 --cpu-n 96 > gpurun_out/f_bench_c3_320.json 2> gpurun_out/f_bench_c3_320.err
timeout 1500 python bench.py --workload c4 --size 512 --steps 20 --warmup 3 --cpu-n 128 > gpurun_out/f_bench_c4_512.json 2> gpurun_out/f_bench_c4_512.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_c3_160.csv python bench.py --workload c3 --size 160 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/f_ncu_launch_c3.log 2>&1
for f in gpurun_out/f_smoke.log gpurun_out/f_pytest.log; do tail -n 4 $f; done
