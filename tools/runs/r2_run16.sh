mkdir -p gpurun_out/r2
timeout 420 python -m pytest tests/test_gpu_ws.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2/pytest_ws16.log
tail -5 gpurun_out/r2/pytest_ws16.log
sw() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --sweep-only 2>>gpurun_out/r2/sweep16.err | tail -1 > gpurun_out/r2/sweep16_$name.json
}
sw ws
sw nows CWN_B200_WS=0
sw ws_dirty CWN_BENCH_DIRTY_FLUSH=1
sw nows_dirty CWN_B200_WS=0 CWN_BENCH_DIRTY_FLUSH=1
sw ws_t64 CWN_B200_WS_TILE=64 CWN_B200_WS_MIN_STAGES=3
sw ws_t32 CWN_B200_WS_TILE=32 CWN_B200_WS_MIN_STAGES=3
sw ws16 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_ws16.so
python - <<'PY'
import json
for f in ('ws','nows','ws_dirty','nows_dirty','ws_t64','ws_t32','ws16'):
    try:
        d=json.loads(open(f'gpurun_out/r2/sweep16_{f}.json').read())
        print(f, ' | '.join(f"{r['kernel'][4:]} {r['adjacency']} {r['F']}: {r['ms']*1e3:.1f}us {r['frac_of_peak']:.3f}" for r in d['kernel_sweep']))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csr_ws_kernel -c 3 -o gpurun_out/r2/ncu_ws16 python bench.py --sweep-only > gpurun_out/r2/ncu_ws16.log 2>&1
tail -3 gpurun_out/r2/ncu_ws16.log
