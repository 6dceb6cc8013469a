#!/bin/bash
# Round 2: (1) N=16 batched convolution, line-ring kernel (default) against the resident-plane kernel at the cell
# counts of the 1/2/4/8-GPU strong-scaling points; (2) ncu launch lists of one shock step at 640 and 80 cells
# (what the fixed, non-convolution part of the step consists of).
mkdir -p gpurun_out
for cells in 640 320 160 80; do
  timeout 60 python tools/gpu_batch_time.py 16 $cells
  SBTE_N16_PLANE=1 timeout 60 python tools/gpu_batch_time.py 16 $cells
done 2>&1 | tee gpurun_out/r02_ring16_ab2.log
for cells in 640 80; do
  SBTE_TOTAL_CELLS=$cells SBTE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_shock_${cells}.csv python bench.py --workload shock_strong --steps 2 --warmup 3 --no-cpu \
    > gpurun_out/r02_launches_shock_${cells}.log 2>&1
done
SBTE_TOTAL_CELLS=80 timeout 120 python bench.py --workload shock_strong --steps 50 --warmup 5 --no-cpu > gpurun_out/r02_shock80_1gpu.json 2>/dev/null
SBTE_TOTAL_CELLS=80 SBTE_NO_GRAPH=1 timeout 120 python bench.py --workload shock_strong --steps 50 --warmup 5 --no-cpu > gpurun_out/r02_shock80_1gpu_nograph.json 2>/dev/null
cut -c1-400 gpurun_out/r02_shock80_1gpu.json gpurun_out/r02_shock80_1gpu_nograph.json
