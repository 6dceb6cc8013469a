#!/bin/bash
# Round 2: final form of the cut-chunk change (whole-chunk instance at N = 16, per-N code shape, builder's choice of cuts)
# against the previous commit's build: parity subset, convolution timings, one-GPU steps.
mkdir -p gpurun_out
P=$PWD/tools/ab/libsbte_b200_prev.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "batched or split_tiles or two_rank or peer_memory or 1d_step or shock1p2 or heattrans or golden" > gpurun_out/r02_cuts6_tests.log 2>&1
tail -2 gpurun_out/r02_cuts6_tests.log
{
for cfg in "16 640" "16 320" "16 160" "16 80" "16 76" "24 250" "24 63" "24 32" "22 250" "22 32" "20 250" "16 640" "24 250"; do
  set -- $cfg
  echo "new : $(timeout 90 python tools/gpu_batch_time.py $1 $2)"
  echo "prev: $(SBTE_LIB_PATH=$P timeout 90 python tools/gpu_batch_time.py $1 $2)"
done
for w in "shock_strong 640" "shock_strong 320" "shock_strong 160" "shock_strong 80" "heattrans_strong 250" "heattrans_strong 63" "heattrans_strong 32" "heattrans22_strong 250" "heattrans22_strong 32"; do
  set -- $w
  for v in new prev; do
    unset SBTE_LIB_PATH
    if [ $v = prev ]; then export SBTE_LIB_PATH=$P; fi
    SBTE_TOTAL_CELLS=$2 timeout 120 python bench.py --workload $1 --steps 40 --warmup 5 --no-cpu > gpurun_out/r02_final_$1_$2_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_final_$1_$2_$v.json'));print('$1 cells=$2 $v', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4))"
  done
done
} 2>&1 | tee gpurun_out/r02_cuts_ab6.log
