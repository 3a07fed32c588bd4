set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z_build.log 2>&1
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "sync_magnetic or c2_3d_pml-200-0 or aniso_smooth or cyl_m1-150 or lorentz_3d-60-3 or known_results-f64" > gpurun_out/z_pytest.log 2>&1; echo "rc=$?"
tail -n 3 gpurun_out/z_pytest.log
