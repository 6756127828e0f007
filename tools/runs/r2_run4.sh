mkdir -p gpurun_out/r2
bash tools/tc5_probe3_run.sh > /dev/null 2>&1
(for a in elu tanh; do for n in 16 64; do
ACT=$a python tools/dbg_grad_err.py $n 0 1 2 3 2>&1 | grep -E "^act"
ACT=$a CWN_B200_DENSE_TC5=0 python tools/dbg_grad_err.py $n 0 1 2 3 2>&1 | grep -E "^act"
done; done) > gpurun_out/r2/grad_err_smooth.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2/pytest_tc5.log
CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_phase.so python tools/phase_timing.py > gpurun_out/r2/phase_tc5.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b_tc5.err | tail -1 > gpurun_out/r2/bench_tc5.json
tail -5 gpurun_out/r2/pytest_tc5.log; cat gpurun_out/r2/grad_err_smooth.txt
