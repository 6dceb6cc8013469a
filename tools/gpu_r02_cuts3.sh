#!/bin/bash
# Round 2: cut chunks, third pass (whole-chunk loop kept apart from the general step loop): new library, default and
# forced modes, against the previous commit's build, interleaved twice; then the one-GPU steps at the full meshes.
mkdir -p gpurun_out
P=$PWD/tools/ab/libsbte_b200_prev.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "batched or split_tiles or two_rank or peer_memory or 1d_step or shock1p2 or heattrans or golden" > gpurun_out/r02_cuts3_tests.log 2>&1
tail -2 gpurun_out/r02_cuts3_tests.log
{
for rep in 1 2; do
  for cfg in "24 250" "16 640" "22 250" "20 250" "24 63" "16 160" "20 63"; do
    set -- $cfg
    echo "default: $(timeout 90 python tools/gpu_batch_time.py $1 $2)"
    echo "any    : $(SBTE_CHUNK_CUTS=0 timeout 90 python tools/gpu_batch_time.py $1 $2)"
    echo "whole  : $(SBTE_CHUNK_CUTS=1 timeout 90 python tools/gpu_batch_time.py $1 $2)"
    echo "prev   : $(SBTE_LIB_PATH=$P timeout 90 python tools/gpu_batch_time.py $1 $2)"
  done
done
for w in "shock_strong 640" "heattrans_strong 250" "heattrans22_strong 250"; do
  set -- $w
  for v in new prev; do
    unset SBTE_LIB_PATH
    if [ $v = prev ]; then export SBTE_LIB_PATH=$P; fi
    SBTE_TOTAL_CELLS=$2 timeout 120 python bench.py --workload $1 --steps 40 --warmup 5 --no-cpu > gpurun_out/r02_full_$1_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_full_$1_$v.json'));print('$1 cells=$2 $v', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4))"
  done
done
} 2>&1 | tee gpurun_out/r02_cuts_ab3.log
