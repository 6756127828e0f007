mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2/pytest_gpu_0.log
CWN_B200_TEST_TC=1 CWN_B200_DENSE_TC=1 python -m pytest tests -m gpu -x -q -k "tensor_core or dense or golden or train" 2>&1 | tail -15 > gpurun_out/r2/pytest_tc.log
python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b0.err | tail -1 > gpurun_out/r2/bench_default.json
CWN_B200_DENSE_TC=1 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b1.err | tail -1 > gpurun_out/r2/bench_tc_mmasync.json
