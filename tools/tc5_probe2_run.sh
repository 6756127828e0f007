#!/bin/bash
mkdir -p gpurun_out/r2
P=tools/_bin/tc5_probe2
O=gpurun_out/r2/tc5_probe2.txt
: > $O
run() { timeout 30 $P "$@" >> $O 2>&1 || echo "FAILED rc=$? args: $*" >> $O; }
run 0 0 64 64 128 0      # baseline (known good)
run 0 0 64 64 64 0       # M = 64 lane mapping
run 0 0 64 128 64 0
for gm in 0 1; do
  run 0 1 64 64 128 $gm  # B MN-major
  run 1 0 64 64 128 $gm  # A MN-major
  run 1 1 64 64 128 $gm
  run 1 1 64 128 128 $gm
  run 1 1 64 64 64 $gm
  run 1 1 32 64 64 $gm
done
cat $O
