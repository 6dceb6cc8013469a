#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for R in 8 4; do
echo "=== heattrans22 R=$R"; SBTE_ANY_R=$R timeout 600 python bench.py --workload heattrans22 --steps 10 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
echo "=== parity R=8"; SBTE_ANY_R=8 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "22 or 12 or any" 2>&1 | tail -2
