#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu (batched + 1d + golden)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "batched or 1d_step or two_rank or heat or regrowth or driver" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_b.log
echo "=== bench 1d v2 640"; timeout 600 python bench.py --workload shock1p2 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'])" | tee gpurun_out/bench_1d_v2b.log
echo "=== bench 1d v2 80"; SBTE_CELLS_PER_GPU=80 timeout 600 python bench.py --workload shock1p2 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'])" | tee gpurun_out/bench_1d_v2b_80.log
echo "=== ncu full batch2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_batch2 -s 2 -c 1 -o gpurun_out/prof_batch2b_r01 python bench.py --workload shock1p2 --steps 1 --warmup 3 > gpurun_out/ncu_full_batch2b.log 2>&1; tail -2 gpurun_out/ncu_full_batch2b.log
