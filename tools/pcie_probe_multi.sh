#!/bin/bash
# Host <-> device copy ceiling with N GPUs copying at once (one process per GPU), 38.5 MB pinned buffers
# like the e2e path of a 256-image call.   tools/pcie_probe_multi.sh <N>
N=${1:-1}
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python tools/pcie_probe.py > gpurun_out/pcie_${N}_$i.log 2>&1 &
done
wait
for i in $(seq 0 $((N-1))); do echo "N=$N gpu $i: $(tr '\n' ' ' < gpurun_out/pcie_${N}_$i.log)"; done
