mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2/pytest_tc5.log
CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_phase.so python tools/phase_timing.py 2>&1 | grep -v Warn > gpurun_out/r2/phase_tc5.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b_tc5.err | tail -1 > gpurun_out/r2/bench_tc5.json
tail -5 gpurun_out/r2/pytest_tc5.log
