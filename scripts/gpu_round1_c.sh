set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/c_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "kernels or f32 or c2_3d or cond_chi3 or lorentz or unfused or known_results or three_d or xperiodic" > gpurun_out/c_pytest.log 2>&1
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/c_bench_512.json 2> gpurun_out/c_bench_512.err
timeout 900 python bench.py --n 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_256.json 2> gpurun_out/c_bench_256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c_launches_256.csv python bench.py --n 256 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/c_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3_kernel -s 6 -c 2 -o gpurun_out/c_prof_step3 python bench.py --n 256 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/c_ncu_full.log 2>&1
for f in gpurun_out/c_smoke.log gpurun_out/c_pytest.log; do tail -n 5 $f; done
