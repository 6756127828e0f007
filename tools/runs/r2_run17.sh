mkdir -p gpurun_out/r2
timeout 420 python -m pytest tests/test_gpu_ws.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2/pytest_ws17.log
tail -5 gpurun_out/r2/pytest_ws17.log
sw() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --sweep-only 2>>gpurun_out/r2/sweep17.err | tail -1 > gpurun_out/r2/sweep17_$name.json
}
sw ws16
sw nows CWN_B200_WS=0
sw ws20 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_ws20.so
sw ws24 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_ws24.so
sw ws16u4 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_ws16u4.so
sw ws16_t64 CWN_B200_WS_TILE=64 CWN_B200_WS_MIN_STAGES=3
python - <<'PY'
import json
for f in ('ws16','nows','ws20','ws24','ws16u4','ws16_t64'):
    try:
        d=json.loads(open(f'gpurun_out/r2/sweep17_{f}.json').read())
        print(f, ' | '.join(f"{r['kernel'][4:]} {r['adjacency'][5:]} {r['F']}: {r['ms']*1e3:.1f} {r['frac_of_peak']:.3f}" for r in d['kernel_sweep']))
    except Exception as e: print(f, 'ERR', e)
PY
