#!/bin/bash
# Round 2: main + split tiles of an N = 16 slab in one launch (SBTE_NO_PAIR_LAUNCH=1: two launches): parity, timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "split_tiles or two_rank or peer_memory or 1d_step or shock1p2 or cuts_inside" > gpurun_out/r02_pair_tests.log 2>&1
tail -2 gpurun_out/r02_pair_tests.log
{
for cells in 80 76 300 44; do
  echo "pair    : $(timeout 60 python tools/gpu_batch_time.py 16 $cells)"
  echo "separate: $(SBTE_NO_PAIR_LAUNCH=1 timeout 60 python tools/gpu_batch_time.py 16 $cells)"
done
for v in 0 1; do
  if [ $v = 1 ]; then export SBTE_NO_PAIR_LAUNCH=1; else unset SBTE_NO_PAIR_LAUNCH; fi
  for cells in 80 76; do
    SBTE_TOTAL_CELLS=$cells timeout 120 python bench.py --workload shock_strong --steps 60 --warmup 5 --no-cpu > gpurun_out/r02_pair_shock${cells}_$v.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02_pair_shock${cells}_$v.json'));print('shock cells=$cells NO_PAIR_LAUNCH=$v', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4), 'launches/step', d['gpu_launches']/d['steps'])"
  done
done
} 2>&1 | tee gpurun_out/r02_pair_ab.log
