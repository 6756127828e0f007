mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2/pytest_r20.log
tail -4 gpurun_out/r2/pytest_r20.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b20.err | tail -1 > gpurun_out/r2/bench_r20_fused.json
CWN_B200_FUSE_REDUCE=0 timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>>gpurun_out/r2/b20.err | tail -1 > gpurun_out/r2/bench_r20_nofuse.json
python - <<'PY'
import json
for f in ('fused','nofuse'):
    try:
        d=json.loads(open(f'gpurun_out/r2/bench_r20_{f}.json').read())
        print(f, round(d['ms_per_step'],4), int(d['value']), int(d['e2e']['value']), d['gpu_launches'], d['config'].get('last_loss'), d.get('ragged',{}).get('ms_per_step'), d.get('ragged',{}).get('eager_ms_per_step'))
        print({k: round(v,4) for k,v in d['roofline']['kernel_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2/b20.err
