#!/bin/bash
# Round 2, final build: compute-sanitizer over the kernels changed last (cut chunks with hand-over, stencil with short
# boundary rows / fused no-flux ends / acquire polling) and fresh ncu evidence for them.
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "split_tiles and 44 or peer_memory or two_rank and 16-8 or 1d_step or batched_computeq and 8-37 or heat_transport_golden" \
  > gpurun_out/r02_memcheck_final.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/r02_memcheck_final.log
tail -4 gpurun_out/r02_memcheck_final.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "split_tiles and 12 or peer_memory and 2-6 or two_rank and 16-8" > gpurun_out/r02_synccheck_final.log 2>&1; echo "synccheck exit $?" | tee -a gpurun_out/r02_synccheck_final.log
tail -3 gpurun_out/r02_synccheck_final.log
# launch list of one 80-cell shock step (no graph, so that every launch is listed)
SBTE_TOTAL_CELLS=80 SBTE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_shock_80_final.csv python bench.py --workload shock_strong --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l80.log 2>&1
# the convolution of a 160-cell slab (cuts inside chunks) and of a 640-cell slab (whole-chunk instance)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring16_cuts_160cells python tools/gpu_batch_time.py 16 160 > gpurun_out/ncu_c160.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 3 -c 1 -f -o gpurun_out/r02_k2_ring16_whole_640cells python tools/gpu_batch_time.py 16 640 > gpurun_out/ncu_w640.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
