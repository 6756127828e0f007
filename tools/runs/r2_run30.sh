mkdir -p gpurun_out/r2
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2/pytest_r30.log; tail -2 gpurun_out/r2/pytest_r30.log
CWN_B200_DENSE_TC5=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "train_step or dense or fused" 2>&1 | tail -2
timeout 200 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), int(d['value']), d['config']['last_loss'])"
