#!/usr/bin/env bash
# single-GPU record run: tests, smoke, the three bench workloads (with cpu_baseline), reference arms
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r01_0d_n32.json | cut -c1-2500
echo "=== bench shock1p2"; timeout 900 python bench.py --workload shock1p2 --steps 20 2>&1 | tail -1 | tee gpurun_out/bench_r01_shock1p2.json | cut -c1-2000
echo "=== bench heattrans"; timeout 900 python bench.py --workload heattrans --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_r01_heattrans.json | cut -c1-2000
echo "=== ref 0d"; timeout 900 python bench.py --impl reference --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_r01_ref_0d.json | cut -c1-700
echo "=== ref 1d"; timeout 900 python bench.py --impl reference --workload shock1p2 --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_r01_ref_shock.json | cut -c1-700
