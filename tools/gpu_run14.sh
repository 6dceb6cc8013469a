#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu (0d)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider -k "bkw or golden or host_driver or reference_driver or smoke or regrowth" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_0d.log
echo "=== bench bkw16"; timeout 900 python bench.py --workload bkw16 --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_bkw16.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_without_graph'], d['e2e']['value'], d['gpu_launches'])"
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
