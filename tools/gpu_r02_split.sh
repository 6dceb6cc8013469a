#!/bin/bash
# Round 2: split tiles for the N = 16 remainder group (SBTE_NO_SPLIT16=1 switches them off): parity, then timing at the
# cell counts where they apply (80 = 8-GPU share of the 640-cell mesh, 75/76 = of the 601-cell mesh, 300 = 2-GPU share).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "split_tiles or batched_computeq or two_rank or peer_memory or 1d_step or shock1p2" > gpurun_out/r02_split_tests.log 2>&1
tail -3 gpurun_out/r02_split_tests.log
for cells in 80 76 300 44 640; do
  timeout 60 python tools/gpu_batch_time.py 16 $cells
  SBTE_NO_SPLIT16=1 timeout 60 python tools/gpu_batch_time.py 16 $cells
done 2>&1 | tee gpurun_out/r02_split_ab.log
for v in 0 1; do
  if [ $v = 1 ]; then export SBTE_NO_SPLIT16=1; fi
  SBTE_TOTAL_CELLS=80 timeout 120 python bench.py --workload shock_strong --steps 50 --warmup 5 --no-cpu > gpurun_out/r02_shock80_split_$v.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02_shock80_split_$v.json'));print('NO_SPLIT16=$v', d['value'], d['ms_per_step'], d['kernel_ms'], d['non_kernel_ms'])"
done
