#!/bin/bash
# A/B of library variants on ONE box (same GPU, back to back): tools/ab_run.sh variant1 variant2 ...
# (variants = libcwn_b200_<name>.so built with cwn_b200.build.build_variant; "default" = the in-tree library)
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), int(d['value']), 'e2e', int(d['e2e']['value']))"; }
for v in "$@"; do
  if [ "$v" = default ]; then run default; else CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_$v.so run $v; fi
done
