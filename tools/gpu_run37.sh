#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_dropin_driver.py -q -x 2>&1 | tail -5 | cut -c1-400
echo "=== quick parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "peer or two_rank or 1d_step or n32 or dropin" 2>&1 | tail -2
