#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_driver.py -q -x 2>&1 | tail -3 | cut -c1-300
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["gpu_launches"], d["e2e"]["checksum"])'
echo "=== shock fused"; timeout 600 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 | python -c "$P"
echo "=== shock unfused"; SBTE_NO_FUSE=1 timeout 600 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 | python -c "$P"
