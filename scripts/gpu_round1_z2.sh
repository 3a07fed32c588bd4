set -x
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z2_build.log 2>&1
timeout 100 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "f32 and (gyro or bfast or aniso_sigma or cond_chi3)" > gpurun_out/z2_pytest.log 2>&1; echo "rc=$?"
tail -n 3 gpurun_out/z2_pytest.log
