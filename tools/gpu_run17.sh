#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/gpu_mp32.py 2>&1 | tail -2
REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 3 -c 1 -o gpurun_out/mp32_stream -f python tools/gpu_mp32.py 2>&1 | tail -3
