mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2/pytest_r9.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b9.err | tail -1 > gpurun_out/r2/bench_r9.json
tail -8 gpurun_out/r2/pytest_r9.log
