#!/bin/bash
# Round 2: small-slab step: cut chunks + march length per slab size + fused no-flux ends + batched epilogue loads,
# against the previous commit's library (tools/ab/libsbte_b200_prev.so) on the same box.  Full parity suite first.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_gpu_tests_g.log 2>&1
tail -3 gpurun_out/r02_gpu_tests_g.log
P=$PWD/tools/ab/libsbte_b200_prev.so
{
for cells in 80 160 320 640; do
  for v in new prev nofuse; do
    unset SBTE_LIB_PATH SBTE_NO_EDGE_FUSE
    if [ $v = prev ]; then export SBTE_LIB_PATH=$P; fi
    if [ $v = nofuse ]; then export SBTE_NO_EDGE_FUSE=1; fi
    SBTE_TOTAL_CELLS=$cells timeout 120 python bench.py --workload shock_strong --steps 60 --warmup 5 --no-cpu > gpurun_out/r02_shock${cells}_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_shock${cells}_$v.json'));print('shock cells=$cells $v', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4))"
  done
done
for cells in 32 63 250; do
  for v in new prev; do
    unset SBTE_LIB_PATH
    if [ $v = prev ]; then export SBTE_LIB_PATH=$P; fi
    SBTE_TOTAL_CELLS=$cells timeout 120 python bench.py --workload heattrans_strong --steps 30 --warmup 5 --no-cpu > gpurun_out/r02_heat${cells}_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_heat${cells}_$v.json'));print('heat cells=$cells $v', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4))"
  done
done
} 2>&1 | tee gpurun_out/r02_smallslab_ab.log
