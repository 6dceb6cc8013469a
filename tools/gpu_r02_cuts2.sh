#!/bin/bash
# Round 2: cut chunks, second pass: the new library (default and SBTE_CHUNK_CUTS=1) against the previous commit's build
# (tools/ab/libsbte_b200_prev.so, SBTE_LIB_PATH), interleaved twice on one box.
mkdir -p gpurun_out
P=$PWD/tools/ab/libsbte_b200_prev.so
{
for rep in 1 2; do
  for cfg in "24 250" "16 640" "22 250" "24 63" "16 160"; do
    set -- $cfg
    echo "new   : $(timeout 90 python tools/gpu_batch_time.py $1 $2)"
    echo "whole : $(SBTE_CHUNK_CUTS=1 timeout 90 python tools/gpu_batch_time.py $1 $2)"
    echo "prev  : $(SBTE_LIB_PATH=$P timeout 90 python tools/gpu_batch_time.py $1 $2)"
  done
done
} 2>&1 | tee gpurun_out/r02_cuts_ab2.log
