#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
: > gpurun_out/scale_r01_8gpu_b.jsonl
run() { echo "=== $*"; env "$1" "$2" timeout 300 $TR bench.py --gpus 8 --workload "$3" --steps "$4" --warmup 5 --no-cpu 2>&1 | tail -1 | tee -a gpurun_out/scale_r01_8gpu_b.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["halo"][:12], d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["e2e"]["checksum"])'; }
run SBTE_HALO=p2p X=1 shock1p2 20
run SBTE_HALO=p2p SBTE_CELLS_PER_GPU=80 shock1p2 50
run SBTE_HALO=nccl SBTE_CELLS_PER_GPU=80 shock1p2 50
run SBTE_HALO=p2p X=1 heattrans 20
