#!/bin/bash
# ncu --set full of one op for several library builds (after a plain run of each).
# Usage: tools/gpu_ncu_ab.sh <op> <batch> <name=path> ...
mkdir -p gpurun_out
op=$1; B=$2; shift; shift
for spec in "$@"; do
  name=${spec%%=*}; path=${spec#*=}
  CHB_LIB=$PWD/$path python tools/op_sweep.py --only $op --batch $B --iters 3 > gpurun_out/plain_${op}_$name.log 2>&1 &&
  CHB_LIB=$PWD/$path ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 2 -c 1 -f -o gpurun_out/prof_${op}_$name \
      python tools/op_sweep.py --only $op --batch $B --iters 3 > gpurun_out/ncu_${op}_$name.log 2>&1
  tail -1 gpurun_out/plain_${op}_$name.log | cut -c1-200
done
