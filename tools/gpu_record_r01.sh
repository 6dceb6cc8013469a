#!/usr/bin/env bash
# Round-1 record run: every bench line, the ncu launch list of the default command, one full capture of the
# dominant kernel, the whole GPU test-suite and smoke().  Outputs land in gpurun_out/ and are copied to profiles/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r01_gpu_tests.log
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_0d_n32.json; cut -c1-600 gpurun_out/bench_r01_0d_n32.json
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 10 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_ref_0d.json; cut -c1-400 gpurun_out/bench_r01_ref_0d.json
for wl in bkw16 shock1p2 heattrans heattrans22; do
  st=20; [ $wl = bkw16 ] && st=100
  echo "=== bench $wl"; timeout 900 python bench.py --workload $wl --steps $st 2>&1 | tail -1 > gpurun_out/bench_r01_$wl.json; cut -c1-300 gpurun_out/bench_r01_$wl.json
done
echo "=== shock 601 cells (the shipped mesh size, prime)"; SBTE_CELLS_PER_GPU=601 timeout 900 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r01_shock601.json; cut -c1-200 gpurun_out/bench_r01_shock601.json
echo "=== ncu launch list (default command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_0d_n32.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1
grep -c "qhat_stream" gpurun_out/r01_launches_bench_0d_n32.csv
echo "=== ncu full: dominant kernel of the default command"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 4 -c 1 -o gpurun_out/r01_k2_stream_sym_final -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b2.log 2>&1; tail -1 gpurun_out/b2.log | cut -c1-100
