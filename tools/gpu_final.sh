#!/bin/bash
# Evidence run for profiles/: headline bench, DRAM traffic of a step, launch list, ncu --set full of the
# headline pass kernel and of single-op batches, per-op sweep at 8192 images (BASELINE.json configs[4]).
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
grep -c "^{" gpurun_out/bench.log
bash tools/gpu_traffic.sh 2>&1 | tail -2
bash tools/gpu_launch_list.sh 2>&1 | tail -16
bash tools/gpu_ncu_mixed.sh 256 2>&1 | tail -1
bash tools/gpu_ncu_ops.sh "Rotate Invert Equalize Sharpness" 2048 2>&1 | tail -4
(timeout 600 python tools/op_sweep.py --batch 8192 --iters 10 --out gpurun_out/op_sweep_8192.json > gpurun_out/sweep8192.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep8192.log)
tail -1 gpurun_out/sweep8192.log
(timeout 300 python bench.py --policy autoaugment --batch 4096 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_autoaugment_4096.log 2>&1)
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1)
