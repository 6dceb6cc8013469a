#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== memcheck: 0D paths + peer halo + 1D step"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "peer_memory or 1d_step or maxpreserve or step_0d or n32 or bkw16 or conserve or moments" 2>&1 | grep -v "^$" | tail -15 | cut -c1-240 | tee gpurun_out/r01_memcheck.log
echo "=== racecheck: cluster kernels (small N)"
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "maxpreserve or step_0d" 2>&1 | grep -v "^$" | tail -12 | cut -c1-240 | tee gpurun_out/r01_racecheck.log
