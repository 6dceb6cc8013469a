#!/bin/bash
# Round 2: ncu --set full of qhat_batch3_kernel<16,0> (640 cells): this build against the previous commit's library.
mkdir -p gpurun_out
P=$PWD/tools/ab/libsbte_b200_prev.so
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring16_new python tools/gpu_batch_time.py 16 640 > gpurun_out/ncu_r16_new.log 2>&1
SBTE_LIB_PATH=$P timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring16_prev python tools/gpu_batch_time.py 16 640 > gpurun_out/ncu_r16_prev.log 2>&1
tail -n 2 gpurun_out/ncu_r16_new.log gpurun_out/ncu_r16_prev.log
