#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== generator timing"; timeout 900 python tools/gpu_gen_time.py 2>&1 | tail -5 | tee gpurun_out/gen_time.log
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-500 | tee gpurun_out/ref_arm.log
