#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== n24"; timeout 900 python tools/gpu_n24.py 2>&1 | tail -5 | tee gpurun_out/n24.log
echo "=== bench 1d real weights"; timeout 600 python bench.py --workload shock1p2 --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1d_real.log | cut -c1-1200
