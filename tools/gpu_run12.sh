#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "=== bench shock"; timeout 900 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_sym_shock2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'])"
echo "=== ncu launches 1d"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1d_sym_r01.csv python bench.py --workload shock1p2 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
echo "=== ncu full stream sym"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_stream -s 3 -c 1 -o gpurun_out/prof_stream_sym_r01 python bench.py --steps 3 --warmup 3 --no-cpu --weights synthetic > gpurun_out/ncu_full_sym.log 2>&1; tail -1 gpurun_out/ncu_full_sym.log
echo "=== ncu full batch2 sym"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:qhat_batch2 -s 2 -c 1 -o gpurun_out/prof_batch2_sym_r01 python bench.py --workload shock1p2 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_b2sym.log 2>&1; tail -1 gpurun_out/ncu_full_b2sym.log
