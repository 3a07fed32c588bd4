set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n_build.log 2>&1
# A/B: one-component-per-thread PML kernel + run-length halos (new defaults) vs the previous forms
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_512_new.json 2> gpurun_out/n_bench_512_new.err
cat gpurun_out/n_bench_512_new.json
MEEP_B200_SPLIT_PML=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_512_nosplit.json 2> gpurun_out/n_bench_512_nosplit.err
cat gpurun_out/n_bench_512_nosplit.json
MEEP_B200_HALO_RUNS=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_512_noruns.json 2> gpurun_out/n_bench_512_noruns.err
cat gpurun_out/n_bench_512_noruns.json
# parity of everything that changed
timeout 1500 python -m pytest tests -m gpu -q -x -k "not reference_test_program and not multigpu" > gpurun_out/n_pytest.log 2>&1
tail -n 3 gpurun_out/n_pytest.log
# ncu: launch list of the default bench command, and a full capture of the fused kernels at 512^3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_c2_512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3 -s 12 -c 4 -o gpurun_out/n_prof_step3_512 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/n_ncu_full.log 2>&1
ls -la gpurun_out/n_*
