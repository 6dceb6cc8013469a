#!/bin/bash
# Round 2: code-shape variants of the cut-chunk kernel (tools/ab/libsbte_b200_{A,B,D}.so) against the previous commit's build.
mkdir -p gpurun_out
{
for cfg in "16 640" "24 250" "16 160" "16 80" "24 63" "22 250" "20 250" "16 640" "24 250"; do
  set -- $cfg
  for v in A B D prev; do
    echo "$v: $(SBTE_LIB_PATH=$PWD/tools/ab/libsbte_b200_$v.so timeout 90 python tools/gpu_batch_time.py $1 $2)"
  done
done
} 2>&1 | tee gpurun_out/r02_cuts_ab4.log
