#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3 | cut -c1-300
for env in "SBTE_X=1" "SBTE_NO_SPLIT=1" "SBTE_X=1" "SBTE_NO_SPLIT=1"; do
echo "=== bench default $env"; env $env timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['plain_kernel']['kernel_ms'])"
done
