#!/bin/bash
# ncu --set full captures of the pass kernel for single ops (after a plain run of the same command).
# Usage: tools/gpu_ncu_ops.sh "Rotate Equalize ..." [batch]
mkdir -p gpurun_out
B=${2:-2048}
for op in $1; do
  python tools/op_sweep.py --only $op --batch $B --iters 3 > gpurun_out/plain_$op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-resident_kernel} -s 2 -c 1 -f -o gpurun_out/prof_op_$op \
      python tools/op_sweep.py --only $op --batch $B --iters 3 > gpurun_out/ncu_$op.log 2>&1
  tail -2 gpurun_out/plain_$op.log | cut -c1-200
done
