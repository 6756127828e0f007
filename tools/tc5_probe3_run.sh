#!/bin/bash
mkdir -p gpurun_out/r2
P=tools/_bin/tc5_probe3
O=gpurun_out/r2/tc5_probe3.txt
: > $O
run() { timeout 30 $P "$@" >> $O 2>&1 || echo "FAILED rc=$? args: $*" >> $O; }
for M in 64 128; do
  for N in 64 128 256; do
    run $M $N 8 1
    run $M $N 24 1
    run $M $N 24 2
    [ $N -le 128 ] && run $M $N 24 4
    run $M $N 48 1
  done
done
run 64 64 1 1
run 128 64 1 1
run 128 256 1 1
cat $O
