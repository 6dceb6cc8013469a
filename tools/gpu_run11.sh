#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "=== bench default (sym)"; timeout 900 python bench.py --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_sym_0d.json | cut -c1-2400
echo "=== bench default (no sym)"; SBTE_NO_SYM=1 timeout 900 python bench.py --no-cpu --weights synthetic 2>&1 | tail -1 | tee gpurun_out/bench_nosym_0d.json | cut -c1-1200
echo "=== bench shock (sym)"; timeout 900 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_sym_shock.json | cut -c1-1500
echo "=== bench heattrans (sym)"; timeout 900 python bench.py --workload heattrans --steps 10 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_sym_heat.json | cut -c1-1500
