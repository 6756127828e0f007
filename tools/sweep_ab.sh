#!/bin/bash
# A/B of the per-adjacency kernel sweep (bench.py --sweep-only) over library variants / env switches on ONE box.
# usage: tools/sweep_ab.sh label[:VAR=VAL[,VAR=VAL]] ...   (label "default" = in-tree library, else libcwn_b200_<label>.so)
sw() {
  python bench.py --sweep-only 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith('{'): continue
    for r in json.loads(l).get('kernel_sweep', []):
        print('$1', r['kernel'], r['adjacency'], r['F'], round(r['ms'] * 1e3, 1), 'us', round(r['frac_of_peak'], 3))
"
}
for spec in "$@"; do
  label=${spec%%:*}; envs=""
  [ "$spec" != "$label" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  lib=${label%%+*}
  if [ "$lib" = default ]; then env $envs bash -c "$(declare -f sw); sw $spec"
  else env $envs CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_$lib.so bash -c "$(declare -f sw); sw $spec"; fi
done
