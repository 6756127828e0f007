# Round-2 verification on one B200: smoke, GPU tests, the bench line as the driver runs it, a 2000-step self check, the
# reference arm, the full config-5 sweep, the ncu launch list of one eager step and the DRAM traffic of the dense kernels.
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py 2>$O/bench_default.err | tail -1 > $O/bench_default.json
timeout 300 python bench.py --steps 2000 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>$O/bench_2000.err | tail -1 > $O/bench_2000steps.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1 > $O/bench_reference.json
timeout 300 python bench.py --config ogb --steps 100 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>$O/bench_ogb.err | tail -1 > $O/bench_ogb.json
timeout 600 python bench.py --sweep-full 2>$O/sweep_full.err > $O/sweep_full.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --mode eager --no-sweep --no-cpu-baseline --no-ragged > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'unit_bwd_tc5|linear_fwd_tc5|unit_bwd_reduce_fast' -s 30 -c 16 -o $O/ncu_dense python bench.py --steps 1 --warmup 1 --mode eager --no-sweep --no-cpu-baseline --no-ragged > $O/ncu_dense.log 2>&1
python - <<'PY'
import json
O='gpurun_out/r2f/'
for f in ('bench_default','bench_2000steps','bench_reference','bench_ogb'):
    try:
        d=json.loads(open(O+f+'.json').read())
        print(f, round(d['ms_per_step'],4), int(d['value']), d.get('e2e',{}).get('value'), d.get('clocks'), (d.get('ragged') or {}).get('ms_per_step'), (d.get('ragged') or {}).get('eager_ms_per_step'), (d.get('e2e_gpu_collation') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
wc -l $O/sweep_full.jsonl $O/launches.csv; tail -2 $O/ncu_dense.log
