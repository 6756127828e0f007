mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ws.py -m gpu -q -k "ogb or train_step or collation or exp" 2>&1 | tail -3
b() { env "$@" timeout 120 python bench.py --config ogb --steps 100 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), int(d['value']), int(d['e2e']['value']), d['config']['last_loss'])"; }
b A=1
b CWN_B200_FUSE_OGB_LOOKUP=0
