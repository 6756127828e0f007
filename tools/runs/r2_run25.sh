mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2/pytest_r25.log
tail -3 gpurun_out/r2/pytest_r25.log
b() { env "$@" timeout 200 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), int(d['value']), int(d['e2e']['value']), d['gpu_launches'], d['config']['last_loss'], round(d['roofline']['kernel_ms_per_step']['csr_plan_build_small'],4))"; }
b A=1
b CWN_B200_PLAN_COUNT=0
b A=2
