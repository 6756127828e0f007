mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2/pytest_r19.log
tail -8 gpurun_out/r2/pytest_r19.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b19_symm.err | tail -1 > gpurun_out/r2/bench_r19_n2_symm.json
CWN_BENCH_DP_SYMM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b19_nccl.err | tail -1 > gpurun_out/r2/bench_r19_n2_nccl.json
timeout 200 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>gpurun_out/r2/b19_n1.err | tail -1 > gpurun_out/r2/bench_r19_n1.json
python - <<'PY'
import json
for f in ('n2_symm','n2_nccl','n1'):
    try:
        d=json.loads(open(f'gpurun_out/r2/bench_r19_{f}.json').read())
        print(f, d['n_gpus'], round(d['ms_per_step'],4), int(d['value']), int(d['e2e']['value']), d['config']['allreduce'][:60], d['config'].get('last_loss'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -i "error\|symmetric\|Traceback" gpurun_out/r2/b19_symm.err | head -5
