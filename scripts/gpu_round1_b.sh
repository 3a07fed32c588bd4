set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/b_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "kernels or f32 or c2_3d or c4_aniso or lorentz or unfused or known_results or three_d" > gpurun_out/b_pytest.log 2>&1
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/b_bench_512.json 2> gpurun_out/b_bench_512.err
timeout 900 python bench.py --n 256 --steps 20 --warmup 3 --cpu-n 128 --prec f32 > gpurun_out/b_bench_256_f32.json 2> gpurun_out/b_bench_256_f32.err
MEEP_B200_FUSE=0 timeout 900 python bench.py --n 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_256_unfused.json 2> gpurun_out/b_bench_256_unfused.err
timeout 900 python bench.py --n 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_256.json 2> gpurun_out/b_bench_256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b_launches_256.csv python bench.py --n 256 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3 -s 12 -c 4 -o gpurun_out/b_prof_step3 python bench.py --n 256 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
for f in gpurun_out/b_smoke.log gpurun_out/b_pytest.log; do tail -n 5 $f; done
