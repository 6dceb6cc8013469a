#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== bench 1d v2"; timeout 600 python bench.py --workload shock1p2 --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_1d_v2.log
echo "=== bench 1d v1"; SBTE_BATCH_V1=1 timeout 600 python bench.py --workload shock1p2 --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_1d_v1.log
echo "=== bench 1d v2 80 cells"; SBTE_CELLS_PER_GPU=80 timeout 600 python bench.py --workload shock1p2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_1d_v2_80.log
echo "=== bench 1d v1 80 cells"; SBTE_BATCH_V1=1 SBTE_CELLS_PER_GPU=80 timeout 600 python bench.py --workload shock1p2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_1d_v1_80.log
echo "=== ncu full batch2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_batch2 -s 2 -c 1 -o gpurun_out/prof_batch2_r01 python bench.py --workload shock1p2 --steps 1 --warmup 3 > gpurun_out/ncu_full_batch2.log 2>&1; tail -2 gpurun_out/ncu_full_batch2.log
echo "=== ncu launches 1d"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1d_r01.csv python bench.py --workload shock1p2 --steps 1 --warmup 3 > gpurun_out/ncu_launch_1d.log 2>&1; tail -1 gpurun_out/ncu_launch_1d.log | cut -c1-200
