mkdir -p gpurun_out/r2
timeout 420 python -m pytest tests/test_gpu_ws.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2/pytest_ws23.log
tail -3 gpurun_out/r2/pytest_ws23.log
sw() { name=$1; shift; env "$@" timeout 200 python bench.py --sweep-only 2>>gpurun_out/r2/sweep23.err | tail -1 > gpurun_out/r2/sweep23_$name.json; }
sw vpl2
sw vpl1 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_vpl1.so
sw nows CWN_B200_WS=0
python - <<'PY'
import json
for f in ('vpl2','vpl1','nows'):
    try:
        d=json.loads(open(f'gpurun_out/r2/sweep23_{f}.json').read())
        print(f, ' | '.join(f"{r['kernel'][4:]} {r['adjacency'][5:]} {r['F']}: {r['ms']*1e3:.1f} {r['frac_of_peak']:.3f}" for r in d['kernel_sweep']))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python bench.py --sweep-full 2>gpurun_out/r2/sweepfull23.err > gpurun_out/r2/sweep_full_r23.jsonl
wc -l gpurun_out/r2/sweep_full_r23.jsonl
