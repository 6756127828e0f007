mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2/pytest_r13.log
tail -12 gpurun_out/r2/pytest_r13.log
