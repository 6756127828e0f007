#!/bin/bash
# runs tools/tc5_probe.cu variants (each in its own process, bounded) -> gpurun_out/r2/tc5_probe.txt
mkdir -p gpurun_out/r2
P=tools/_bin/tc5_probe
O=gpurun_out/r2/tc5_probe.txt
: > $O
run() { timeout 30 $P "$@" >> $O 2>&1 || echo "FAILED rc=$? args: $*" >> $O; }
# descriptor semantics: all four major combinations, K = N = 64, single accumulator; then swapped LBO/SBO (must be wrong)
for a in 0 1; do for b in 0 1; do run $a $b 64 64 0 1 0 0; done; done
for a in 0 1; do for b in 0 1; do run $a $b 64 64 0 1 1 0; done; done
# shapes the dense kernels use
run 0 0 128 64 0 1 0 0
run 0 1 64 128 0 1 0 0
run 1 1 128 64 0 1 0 0
run 1 1 128 128 0 1 0 0
# accumulation quality: modes, normal and all-positive data
for pos in 0 1; do
  run 0 0 64 64 3 1 0 $pos
  run 0 0 64 64 0 1 0 $pos
  run 0 0 64 64 1 1 0 $pos
  run 0 0 64 64 2 2 0 $pos
  run 0 0 64 64 2 4 0 $pos
  run 0 0 64 64 2 7 0 $pos
  run 0 0 128 64 0 1 0 $pos
  run 0 0 128 64 1 1 0 $pos
  run 0 0 128 64 2 4 0 $pos
  run 0 0 128 64 2 7 0 $pos
done
cat $O
