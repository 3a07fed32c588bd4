set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/o_build.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_512_split.json 2> gpurun_out/o_bench_512_split.err
cat gpurun_out/o_bench_512_split.json
MEEP_B200_SPLIT_PML=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_512_nosplit.json 2> gpurun_out/o_bench_512_nosplit.err
cat gpurun_out/o_bench_512_nosplit.json
timeout 600 python bench.py --workload c4 --size 512 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_c4_512.json 2> gpurun_out/o_bench_c4_512.err
cat gpurun_out/o_bench_c4_512.json
timeout 600 python bench.py --workload c3 --size 320 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_c3_320.json 2> gpurun_out/o_bench_c3_320.err
cat gpurun_out/o_bench_c3_320.json
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "not reference_test_program" > gpurun_out/o_pytest.log 2>&1
tail -n 3 gpurun_out/o_pytest.log
