#!/bin/bash
# Round 2: ncu --set full of qhat_batch3_kernel<24> (250 cells) with cut chunks (default) and with whole-chunk cuts.
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring24_cuts python tools/gpu_batch_time.py 24 250 > gpurun_out/ncu_cuts.log 2>&1
SBTE_CHUNK_CUTS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring24_whole python tools/gpu_batch_time.py 24 250 > gpurun_out/ncu_whole.log 2>&1
tail -2 gpurun_out/ncu_cuts.log gpurun_out/ncu_whole.log
