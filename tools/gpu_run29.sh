#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_driver.py -q -x 2>&1 | tail -4 | cut -c1-300
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["launch_by_launch_ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["e2e"]["checksum"])'
for wl in shock1p2 heattrans; do
echo "=== $wl graph"; timeout 600 python bench.py --workload $wl --steps 20 --no-cpu 2>&1 | tail -1 | python -c "$P"
done
echo "=== shock 80 cells graph"; SBTE_CELLS_PER_GPU=80 timeout 600 python bench.py --workload shock1p2 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "$P"
echo "=== shock 80 cells no graph"; SBTE_NO_GRAPH=1 SBTE_CELLS_PER_GPU=80 timeout 600 python bench.py --workload shock1p2 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "$P"
