#!/usr/bin/env bash
# Round-2 record run (after tools/gpu_r02_mirror.sh has decided which kernels become the default): smoke(), the whole
# GPU test-suite, every bench line, the reference arm, the ncu launch list of the default command and full captures
# of the dominant kernels.  Outputs land in gpurun_out/; copy what should be judged to profiles/ (r02_ prefix).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests.log
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r02_0d_n32.json; cut -c1-600 gpurun_out/bench_r02_0d_n32.json
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r02_ref_0d.json
for wl in bkw16 shock1p2 heattrans heattrans22; do
  st=20; [ $wl = bkw16 ] && st=100
  echo "=== bench $wl"; timeout 900 python bench.py --workload $wl --steps $st 2>&1 | tail -1 > gpurun_out/bench_r02_$wl.json; cut -c1-300 gpurun_out/bench_r02_$wl.json
done
echo "=== ncu launch list (default command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_0d_n32.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1
echo "=== ncu full: 0D stream kernel, 1D batched kernels (N=16 shock, N=24 heatTrans)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 4 -c 1 -o gpurun_out/r02_k2_stream_n32 -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"qhat_(batch|mirror)" -s 2 -c 1 -o gpurun_out/r02_k2_batched_n16 -f python tools/gpu_n22_time.py 16 640 > gpurun_out/b3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"qhat_(batch|mirror)" -s 2 -c 1 -o gpurun_out/r02_k2_batched_n24 -f python tools/gpu_n22_time.py 24 250 > gpurun_out/b4.log 2>&1
ls -la gpurun_out | tail -20
