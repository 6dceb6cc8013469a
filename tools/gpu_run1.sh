#!/usr/bin/env bash
# First GPU session: parity tests, smoke, bench (both stream depths), ncu launch list + full capture.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench auto"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_auto.log
echo "=== bench deep"; timeout 600 python bench.py --steps 30 --warmup 5 --k2 deep --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_deep.log
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_stream -s 3 -c 2 -o gpurun_out/prof_stream_r01 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
