#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for sp in 2 3 4 8; do
echo "=== bench default split $sp"; SBTE_SPLIT=$sp timeout 600 python bench.py --no-cpu --weights synthetic 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['plain_kernel']['kernel_ms'], d['e2e']['checksum'])"
done
