#!/bin/bash
# ncu --set full of the level-0 pass kernel of a mixed RandAugment(2,10) batch (after a plain run).
mkdir -p gpurun_out
B=${1:-2048}
python tools/op_sweep.py --only RandAugment --batch $B --iters 3 > gpurun_out/plain_mixed.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 3 -c 1 -f -o gpurun_out/prof_mixed \
    python tools/op_sweep.py --only RandAugment --batch $B --iters 3 > gpurun_out/ncu_mixed.log 2>&1
tail -1 gpurun_out/plain_mixed.log | cut -c1-200
