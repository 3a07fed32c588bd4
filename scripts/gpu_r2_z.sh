# round 2, call Z (1 GPU): ncu --set full of the final kernels: fast path in single precision (what bounds it?),
# PML kernel (double, 512^3), Lorentz / f_minus_p / E update on the Au sphere.  Text exports only (64 MiB limit).
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z_build.log 2>&1; tail -n 2 gpurun_out/z_build.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step3_plain -s 10 -c 2 -o /tmp/z_plain_f32 python bench.py --n 512 --prec f32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu_plain_f32.log 2>&1
ncu -i /tmp/z_plain_f32.ncu-rep --page details > gpurun_out/z_plain_512_f32_ncu_details.txt
ncu -i /tmp/z_plain_f32.ncu-rep --page source --csv --print-source sass > gpurun_out/z_plain_512_f32_ncu_sass.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step3c -s 10 -c 2 -o /tmp/z_pml_f64 python bench.py --n 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu_pml_f64.log 2>&1
ncu -i /tmp/z_pml_f64.ncu-rep --page details > gpurun_out/z_step3c_512_ncu_details.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lorentz_blocked|fmp_kernel|edhb_kernel" -s 18 -c 3 -o /tmp/z_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu_c3.log 2>&1
ncu -i /tmp/z_c3.ncu-rep --page details > gpurun_out/z_c3_pols_ncu_details.txt
ls -la gpurun_out/
