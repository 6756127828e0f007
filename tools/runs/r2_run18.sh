mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2/pytest_r18.log
tail -4 gpurun_out/r2/pytest_r18.log
timeout 600 python bench.py --steps 200 --warmup 10 2>gpurun_out/r2/b18.err | tail -1 > gpurun_out/r2/bench_r18.json
timeout 600 python bench.py --sweep-full 2>gpurun_out/r2/sweepfull18.err > gpurun_out/r2/sweep_full_r18.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csr_ws_kernel -c 8 -o gpurun_out/r2/ncu_ws18 python bench.py --sweep-only > gpurun_out/r2/ncu_ws18.log 2>&1
tail -2 gpurun_out/r2/ncu_ws18.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_r18.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('ragged',{}).get('ms_per_step'), d.get('e2e_gpu_collation',{}).get('value'))
for r in d['kernel_sweep']: print(r['kernel'], r['adjacency'], r['F'], round(r['ms']*1e3,1), round(r['frac_of_peak'],3))
print(d['cpu_baseline'])
PY
wc -l gpurun_out/r2/sweep_full_r18.jsonl
