mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2/pytest_r7.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b7.err | tail -1 > gpurun_out/r2/bench_r7.json
timeout 600 python bench.py --config ogb --steps 100 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b7_ogb.err | tail -1 > gpurun_out/r2/bench_r7_ogb.json
tail -8 gpurun_out/r2/pytest_r7.log; tail -3 gpurun_out/r2/b7.err; tail -3 gpurun_out/r2/b7_ogb.err
