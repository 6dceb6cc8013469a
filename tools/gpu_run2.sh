#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== debug1"; timeout 600 python tools/gpu_debug1.py 2>&1 | tail -20 | tee gpurun_out/debug1.log
echo "=== sanitizer bench"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --steps 1 --warmup 3 --no-cpu 2>&1 | grep -v "^$" | head -60 | tee gpurun_out/sanitizer_bench.log
