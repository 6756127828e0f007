mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "train_step or fused or dense or graph or padded or batch" 2>&1 | tail -4 > gpurun_out/r2/pytest_r21.log
tail -2 gpurun_out/r2/pytest_r21.log
b() { env "$@" timeout 200 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), int(d['value']), d['gpu_launches'])"; }
b A=1
b CWN_B200_FUSE_REDUCE=0
b A=2
b CWN_B200_FUSE_REDUCE=0
b A=3
