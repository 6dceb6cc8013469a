#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/gpu_ffttime.py
SBTE_NO_CLUSTER_FFT=1 timeout 300 python tools/gpu_ffttime.py
REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fft3d_cluster|maxwellian" -s 4 -c 3 -o gpurun_out/fft_cluster2 -f python tools/gpu_mp32.py 2>&1 | tail -2
