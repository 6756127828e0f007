mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_f64.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2/pytest_f64_r22.log
tail -30 gpurun_out/r2/pytest_f64_r22.log
