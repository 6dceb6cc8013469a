#!/usr/bin/env bash
# 2-GPU session: torchrun both workloads, N=1 lines on the same box for comparison
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
echo "=== 1d N=1"; timeout 600 python bench.py --workload shock1p2 --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_1d_n1.log | cut -c1-400
echo "=== 1d N=2"; timeout 600 $TR --nproc-per-node 2 bench.py --workload shock1p2 --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/scale_1d_n2.log | cut -c1-400
echo "=== 1d N=2 strong (320 cells/GPU)"; SBTE_CELLS_PER_GPU=320 timeout 600 $TR --nproc-per-node 2 bench.py --workload shock1p2 --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/scale_1d_n2_strong.log | cut -c1-400
echo "=== 0d N=2"; timeout 600 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/scale_0d_n2.log | cut -c1-600
echo "=== ref arm N=2"; timeout 600 $TR --nproc-per-node 2 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 2>&1 | tail -2 | tee gpurun_out/ref_n2.log | cut -c1-300
