#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
echo "=== default bench, 2 GPUs (driver command)"; timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-700
echo "=== reference arm, 2 ranks"; timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
