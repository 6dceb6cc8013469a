#!/bin/bash
# Round 2: stream-K cuts inside a xi_x chunk (line-ring kernels; SBTE_CHUNK_CUTS=1 restores whole-chunk cuts).
# Parity first, then the convolution at the per-GPU cell counts of the strong-scaling points, then the 80-/160-cell steps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "batched or split_tiles or two_rank or peer_memory or 1d_step or shock1p2 or heattrans or invariance or golden" > gpurun_out/r02_cuts_tests.log 2>&1
tail -3 gpurun_out/r02_cuts_tests.log
{
for cells in 640 320 160 80 76 32; do
  timeout 60 python tools/gpu_batch_time.py 16 $cells
  SBTE_CHUNK_CUTS=1 timeout 60 python tools/gpu_batch_time.py 16 $cells
done
for cells in 250 125 63 32; do
  timeout 90 python tools/gpu_batch_time.py 24 $cells
  SBTE_CHUNK_CUTS=1 timeout 90 python tools/gpu_batch_time.py 24 $cells
done
for cells in 250 32; do
  timeout 90 python tools/gpu_batch_time.py 22 $cells
  SBTE_CHUNK_CUTS=1 timeout 90 python tools/gpu_batch_time.py 22 $cells
done
} 2>&1 | tee gpurun_out/r02_cuts_ab.log
for cells in 80 160; do
  for v in 0 1; do
    if [ $v = 1 ]; then export SBTE_CHUNK_CUTS=1; else unset SBTE_CHUNK_CUTS; fi
    SBTE_TOTAL_CELLS=$cells timeout 120 python bench.py --workload shock_strong --steps 50 --warmup 5 --no-cpu > gpurun_out/r02_shock${cells}_cuts_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_shock${cells}_cuts_$v.json'));print('cells=$cells CHUNK_CUTS=$v', d['value'], d['ms_per_step'], d['kernel_ms'], d['non_kernel_ms'])"
  done
done 2>&1 | tee -a gpurun_out/r02_cuts_ab.log
unset SBTE_CHUNK_CUTS
SBTE_TOTAL_CELLS=32 timeout 120 python bench.py --workload heattrans_strong --steps 30 --warmup 5 --no-cpu 2>/dev/null | cut -c1-300
