mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z3_build.log 2>&1
timeout 60 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiled or midrun or phase_in or bloch_change" > gpurun_out/z3_pytest.log 2>&1; echo "rc=$?"
tail -n 2 gpurun_out/z3_pytest.log
