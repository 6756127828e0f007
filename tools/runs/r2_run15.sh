mkdir -p gpurun_out/r2
timeout 420 python -m pytest tests/test_gpu_ws.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2/pytest_ws15.log
timeout 200 python bench.py --sweep-only 2>gpurun_out/r2/sweep15.err | tail -1 > gpurun_out/r2/sweep15_ws.json
CWN_B200_WS=0 timeout 200 python bench.py --sweep-only 2>>gpurun_out/r2/sweep15.err | tail -1 > gpurun_out/r2/sweep15_nows.json
tail -15 gpurun_out/r2/pytest_ws15.log
python - <<'PY'
import json
for f in ('sweep15_ws','sweep15_nows'):
    try:
        d=json.loads(open(f'gpurun_out/r2/{f}.json').read())
        for r in d['kernel_sweep']: print(f, r['kernel'], r['adjacency'], r['F'], round(r['ms']*1e3,1), 'us', round(r['frac_of_peak'],3))
    except Exception as e: print(f, 'ERR', e)
PY
