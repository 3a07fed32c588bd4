# round 2, call A: regression of the changed parity cases, first-step timeline at 512^3, ncu of the
# two kernels VERDICT asked evidence for (halo_kernel, step3c_kernel) on the round-1 kernels
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1; tail -n 2 gpurun_out/a_build.log
timeout 900 python -m pytest tests -m gpu -q -x -k "c3_au or phase_in or midrun or c2_3d_pml or kernels or abi" > gpurun_out/a_pytest.log 2>&1
tail -n 5 gpurun_out/a_pytest.log
MEEP_B200_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_512.json 2> gpurun_out/a_bench_512.err
grep -v "no E/H fusion" gpurun_out/a_bench_512.err | cut -c1-220 | tail -n 80
cat gpurun_out/a_bench_512.json | cut -c1-1500
timeout 900 ncu --set full --clock-control none --import-source on -k regex:halo_kernel -s 8 -c 2 -o gpurun_out/a_prof_halo_512 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/a_ncu_halo.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3c -s 8 -c 2 -o gpurun_out/a_prof_step3c_512 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/a_ncu_step3c.log 2>&1
ls -la gpurun_out/a_*
nproc; free -g | head -2
