#!/bin/bash
# Round-2 profiling pass (one GPU): launch lists of the default bench command and of the sharded shock case, and one
# ncu --set full capture of each dominant kernel.  Summaries are made offline with tools/ncu_summary.py.
mkdir -p gpurun_out
# launch list of the default command (0D legs + the 1D sub-records); numbers printed under ncu are not bench values
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench_default.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-dropin --no-sustained > gpurun_out/r02_launches_bench_default.log 2>&1
# full captures
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 6 -c 1 -o gpurun_out/r02_k2_stream_sym_n32 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-dropin --no-sustained --no-oned > gpurun_out/ncu_stream.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"qhat_batch3_kernel<\(int\)16, \(bool\)0>" -s 2 -c 1 -o gpurun_out/r02_k2_ring16_640cells -f \
  python tools/gpu_batch_time.py 16 640 > gpurun_out/ncu_ring16.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"qhat_batch3_kernel" -s 4 -c 2 -o gpurun_out/r02_k2_ring16_split_80cells -f \
  python tools/gpu_batch_time.py 16 80 > gpurun_out/ncu_ring16_80.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"qhat_batch3_kernel" -s 2 -c 1 -o gpurun_out/r02_k2_ring24_250cells -f \
  python tools/gpu_batch_time.py 24 250 > gpurun_out/ncu_ring24.log 2>&1
SBTE_TOTAL_CELLS=80 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fft3d_cell_kernel|upwind_two_kernel" -s 12 -c 6 -o gpurun_out/r02_small_kernels_80cells -f \
  python bench.py --workload shock_strong --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out/*.ncu-rep
