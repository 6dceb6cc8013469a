#!/usr/bin/env bash
# compute-sanitizer over the small-N parity tests (every kernel family; the N=32 split path is forced onto N=16)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SBTE_SPLIT_ALL=1 timeout 1200 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "not n32 and not bkw16 and not real_weights and not generated" 2>&1 | grep -v "^$" | tail -12 | cut -c1-240 | tee gpurun_out/r01_memcheck.log
