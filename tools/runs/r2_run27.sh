mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2/pytest_r27.log
tail -2 gpurun_out/r2/pytest_r27.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b27_symm.err | tail -1 > gpurun_out/r2/bench_r27_n2_symm.json
CWN_BENCH_DP_SYMM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b27_nccl.err | tail -1 > gpurun_out/r2/bench_r27_n2_nccl.json
timeout 200 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>/dev/null | tail -1 > gpurun_out/r2/bench_r27_n1.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2/bench_r27_ref_n2.json
python - <<'PY'
import json
for f in ('n2_symm','n2_nccl','n1','ref_n2'):
    try:
        d=json.loads(open(f'gpurun_out/r2/bench_r27_{f}.json').read())
        print(f, d['n_gpus'], round(d['ms_per_step'],4), int(d['value']), int(d['e2e']['value']), str(d['config'].get('allreduce'))[:50], d.get('cpu_baseline',{}).get('cores'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -i "symmetric\|Traceback\|Error" gpurun_out/r2/b27_symm.err | head -5
