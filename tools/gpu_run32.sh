#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3 | cut -c1-300
echo "=== heattrans22"; timeout 600 python bench.py --workload heattrans22 --steps 10 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r01_heattrans22.json; python -c "import sys,json; d=json.load(open('gpurun_out/bench_r01_heattrans22.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
