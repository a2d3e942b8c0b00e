#!/bin/bash
# The multi-GPU visit (gpurun --gpus 8): the multi-device parity test, the concurrent PCIe ceiling, weak scaling of
# the headline config and strong scaling of BASELINE configs[2] (AutoAugment, fixed 4096-image batch).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 500 -p no:cacheprovider -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log); tail -3 gpurun_out/pytest_multi.log
bash tools/pcie_probe_multi.sh 1
bash tools/pcie_probe_multi.sh 8
for n in 1 2 4 8; do
  (timeout 300 python bench.py --gpus $n --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/weak_$n.log 2>&1; echo "exit $?" >> gpurun_out/weak_$n.log)
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
for line in open("gpurun_out/weak_%s.log" % n):
    if line.startswith("{"):
        d = json.loads(line)
        print("WEAK   N=%s value %.0f img/s ms/step %.4f frac %.3f e2e %.0f img/s" % (n, d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value", 0)))
PY
done
for n in 1 2 4 8; do
  (timeout 300 python bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline --policy autoaugment --batch 4096 --strong > gpurun_out/strong_$n.log 2>&1; echo "exit $?" >> gpurun_out/strong_$n.log)
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
for line in open("gpurun_out/strong_%s.log" % n):
    if line.startswith("{"):
        d = json.loads(line)
        print("STRONG N=%s value %.0f img/s ms/step %.4f frac(per GPU) %.3f e2e %.0f checksum %s" % (n, d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value", 0), d.get("output_checksum_call0")))
PY
done
grep -h "exit" gpurun_out/weak_*.log gpurun_out/strong_*.log | sort | uniq -c
