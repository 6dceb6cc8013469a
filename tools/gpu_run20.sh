#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -6 | cut -c1-300
echo "=== mp32"; timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
echo "=== bkw16"; timeout 600 python bench.py --workload bkw16 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fft3d_cluster|maxwellian" -s 4 -c 3 -o gpurun_out/fft_cluster -f python tools/gpu_mp32.py 2>&1 | tail -2
