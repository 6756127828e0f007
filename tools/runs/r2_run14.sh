mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2/pytest_r14.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>gpurun_out/r2/b14.err | tail -1 > gpurun_out/r2/bench_r14_pdl.json
CWN_B200_PDL=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>gpurun_out/r2/b14b.err | tail -1 > gpurun_out/r2/bench_r14_nopdl.json
tail -6 gpurun_out/r2/pytest_r14.log
