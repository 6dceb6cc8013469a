#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4 | cut -c1-300
for env in "SBTE_X=1" "SBTE_MP_NARROW=1"; do
echo "=== mp32 $env"; env $env timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
echo "=== bkw16 $env"; env $env timeout 600 python bench.py --workload bkw16 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 3 -c 1 -o gpurun_out/mp32_stream_wide -f python tools/gpu_mp32.py 2>&1 | tail -1
