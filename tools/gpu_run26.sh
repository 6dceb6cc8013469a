#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fft3d_cell_kernel|conserve_kernel|upwind_two" -s 6 -c 4 -o gpurun_out/shock_small -f python bench.py --workload shock1p2 --steps 1 --warmup 3 --no-cpu 2>&1 | tail -2
