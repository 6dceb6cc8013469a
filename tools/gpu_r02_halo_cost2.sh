#!/bin/bash
# Round 2: halo cost after a protocol change: cross-GPU parity first, then the same slab size per GPU on 1 and on 2 GPUs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider -k "two_gpus or peer or two_rank or invariance or 1d_step or shock1p2 or upwind or transport or golden or driver" > gpurun_out/r02_halo2_tests.log 2>&1
tail -3 gpurun_out/r02_halo2_tests.log
for per in 80 160 320; do
  SBTE_TOTAL_CELLS=$per timeout 120 python bench.py --workload shock_strong --steps 60 --warmup 5 --no-cpu > gpurun_out/r02_halo_${per}_1gpu.json 2>/dev/null
  SBTE_TOTAL_CELLS=$((2*per)) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload shock_strong --steps 60 --warmup 5 --no-cpu > gpurun_out/r02_halo_${per}_2gpu.json 2>/dev/null
  python - <<PY
import json
for n in (1,2):
    d=json.loads([l for l in open('gpurun_out/r02_halo_${per}_%dgpu.json'%n) if l.startswith('{')][-1])
    print('cells/GPU=$per gpus=%d'%n, round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['kernel_ms'],4), 'rest', round(d['non_kernel_ms'],4), d.get('halo_parity',''))
PY
done 2>&1 | tee gpurun_out/r02_halo_cost2.log
