set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/w_build.log 2>&1
timeout 400 python bench.py --prec f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_512_f32.json 2> gpurun_out/w_bench_512_f32.err; echo "rc=$?"
cat gpurun_out/w_bench_512_f32.json
