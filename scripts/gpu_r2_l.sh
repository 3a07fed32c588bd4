# round 2, call L (1 GPU): sub-volume readers on the device, c3 halo volume, final-form ncu captures
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/l_build.log 2>&1; tail -n 2 gpurun_out/l_build.log
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels or flux or c2_3d_pml or 3d_metal or dft_fields or sync_magnetic or midrun or stress_tensor or integrate or near2far" > gpurun_out/l_pytest.log 2>&1
tail -n 4 gpurun_out/l_pytest.log
MEEP_B200_VERBOSE=1 timeout 600 python bench.py --workload c3 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/l_bench_c3.json 2> gpurun_out/l_bench_c3.err
grep "recorded step_boundaries\|recorded update_dfts" gpurun_out/l_bench_c3.err | tail -n 12 | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_bench_c3.json').read().strip().splitlines()[-1])
print('c3', round(d['ms_per_step'],3), {k:(v['launches_per_step'], round(v['ms_per_step'],3), round(v['alg_bytes_per_step']/1e6,1)) for k,v in d['roofline']['kernels'].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3c -s 8 -c 2 -o gpurun_out/l_prof_step3c_512 python bench.py --size 512 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/l_ncu_step3c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:halo_kernel -s 8 -c 2 -o gpurun_out/l_prof_halo_512 python bench.py --size 512 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/l_ncu_halo.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:dft_kernel -c 2 -o gpurun_out/l_prof_dft_c3 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/l_ncu_dft.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:flux_ -c 2 -o gpurun_out/l_prof_flux python -m pytest tests/test_kernels_gpu.py -q -k "flux" > gpurun_out/l_ncu_flux.log 2>&1
ls -la gpurun_out/l_*
