mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2/pytest_tc5.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b_tc5.err | tail -1 > gpurun_out/r2/bench_tc5.json
CWN_B200_DENSE_TC5=0 timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b_ffma.err | tail -1 > gpurun_out/r2/bench_ffma.json
tail -5 gpurun_out/r2/pytest_tc5.log; tail -3 gpurun_out/r2/b_tc5.err
