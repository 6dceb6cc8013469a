#!/bin/bash
# Round 2: code-shape variants E, F of the cut-chunk kernel against the previous commit's build.
mkdir -p gpurun_out
{
for cfg in "16 640" "16 160" "16 80" "16 320" "24 250" "16 640" "16 80"; do
  set -- $cfg
  for v in E F prev; do
    echo "$v: $(SBTE_LIB_PATH=$PWD/tools/ab/libsbte_b200_$v.so timeout 90 python tools/gpu_batch_time.py $1 $2)"
  done
done
} 2>&1 | tee gpurun_out/r02_cuts_ab5.log
