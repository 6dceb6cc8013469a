#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "=== bench heattrans22"; timeout 900 python bench.py --workload heattrans22 --steps 5 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_heat22.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
echo "=== bench bkw16"; timeout 900 python bench.py --workload bkw16 --steps 50 2>&1 | tail -1 | tee gpurun_out/bench_bkw16.json | cut -c1-1500
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r01_0d_n32.json | cut -c1-3000
