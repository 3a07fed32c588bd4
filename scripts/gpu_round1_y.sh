set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/y_build.log 2>&1
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "ld_preload or c2_3d_pml-200-0" > gpurun_out/y_pytest.log 2>&1; echo "rc=$?"
tail -n 5 gpurun_out/y_pytest.log
