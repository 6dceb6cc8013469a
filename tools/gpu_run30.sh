#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests (1 GPU)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "peer_memory or two_rank or 1d_step" 2>&1 | tail -3 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["halo"][:12], d["value"], d["ms_per_step"], d["roofline"]["launch_by_launch_ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["e2e"]["checksum"])'
for mode in nccl p2p; do
  echo "=== shock $mode"; SBTE_HALO=$mode timeout 300 $TR bench.py --gpus 2 --workload shock1p2 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P"
  echo "=== shock80 $mode"; SBTE_CELLS_PER_GPU=80 SBTE_HALO=$mode timeout 300 $TR bench.py --gpus 2 --workload shock1p2 --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "$P"
  echo "=== heattrans $mode"; SBTE_HALO=$mode timeout 300 $TR bench.py --gpus 2 --workload heattrans --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P"
done
