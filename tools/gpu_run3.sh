#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench auto"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_auto.log
echo "=== bench deep"; timeout 600 python bench.py --steps 30 --warmup 5 --k2 deep --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_deep.log
echo "=== bench 1d"; timeout 600 python bench.py --workload shock1p2 --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_1d.log
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full stream"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_stream -s 3 -c 2 -o gpurun_out/prof_stream_r01 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== ncu full batch"; SBTE_CELLS_PER_GPU=320 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_batch -s 2 -c 1 -o gpurun_out/prof_batch_r01 python bench.py --workload shock1p2 --steps 1 --warmup 3 > gpurun_out/ncu_full_batch.log 2>&1; tail -2 gpurun_out/ncu_full_batch.log
ls -la gpurun_out
