set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/v_build.log 2>&1
timeout 400 python -m pytest tests -m gpu -q -x -k "noisy or gyro or halo_zero or aniso_sigma or golden" > gpurun_out/v_pytest_new.log 2>&1; echo "new rc=$?"
tail -n 3 gpurun_out/v_pytest_new.log
