mkdir -p gpurun_out/r2
(for n in 16 64; do
python tools/dbg_grad_err.py $n 0 1 2 3
CWN_B200_DENSE_TC5=0 python tools/dbg_grad_err.py $n 0 1 2 3
done) > gpurun_out/r2/grad_err.txt 2>&1
CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_phase.so python tools/phase_timing.py > gpurun_out/r2/phase_tc5.txt 2>&1
cat gpurun_out/r2/grad_err.txt | grep -v Warn
