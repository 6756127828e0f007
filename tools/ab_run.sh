#!/bin/bash
# A/B of library / launch variants on ONE box (same GPU, back to back), then a CUPTI timeline of one replayed step.
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), int(d['value']))"; }
run default
CWN_BENCH_SMI_MS=200 run smi200
CWN_B200_TILE_ROWS=64 CWN_BENCH_SMI_MS=200 run tile64_smi200
for v in nopipe nopipe2 pipe1; do
  CWN_BENCH_SMI_MS=200 CWN_B200_TILE_ROWS=64 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_$v.so run ${v}_tile64
done
CWN_BENCH_SMI_MS=200 CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_nopipe.so run nopipe_tile32
CWN_BENCH_SMI_MS=200 CWN_B200_STREAMS=0 run nostreams
timeout 200 python tools/trace_step.py > gpurun_out/trace_step.txt 2>&1
head -50 gpurun_out/trace_step.txt
