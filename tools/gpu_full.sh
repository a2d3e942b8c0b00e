#!/bin/bash
# Full single-GPU visit: every GPU test, smoke, the default bench (both arms), the sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
(timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -5 gpurun_out/pytest.log | cut -c1-300
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest.log | head -20 | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log); tail -2 gpurun_out/bench.log | cut -c1-1500
(timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref.log); tail -2 gpurun_out/bench_ref.log | cut -c1-700
(timeout 400 python tools/op_sweep.py --batch ${1:-8192} --iters 10 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log)
tail -1 gpurun_out/sweep.log
