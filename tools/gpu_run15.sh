#!/usr/bin/env bash
# peer-memory halo: single-GPU emulation test, then 2-GPU NCCL vs peer-memory runs of both 1D workloads
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== peer test"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "peer_memory or two_rank or 1d_step" 2>&1 | tail -5 | tee gpurun_out/peer_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for wl in shock1p2 heattrans; do
  for mode in nccl p2p; do
    echo "=== $wl $mode"
    SBTE_HALO=$mode timeout 300 $TR bench.py --gpus 2 --workload $wl --steps 20 --warmup 3 --no-cpu 2>&1 | grep -v "^$" | tail -3 | tee gpurun_out/peer_${wl}_${mode}.log
  done
done
echo "=== strong shock 80 cells/GPU"
for mode in nccl p2p; do
  SBTE_CELLS_PER_GPU=80 SBTE_HALO=$mode timeout 300 $TR bench.py --gpus 2 --workload shock1p2 --steps 50 --warmup 5 --no-cpu 2>&1 | grep -v "^$" | tail -2 | tee gpurun_out/peer_shock80_${mode}.log
done
