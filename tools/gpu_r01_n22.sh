#!/bin/bash
# One short GPU call (no torch import before the last step): parity of the line-ring kernel with partial row-blocks
# (N = 20, 22), its timing against the any-N kernel, the N = 24 / 22 cases of the existing suite, then the bench line.
mkdir -p gpurun_out
T="timeout"
$T 90 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "line_ring" > gpurun_out/n22_tests_a.log 2>&1; echo "exit $?" >> gpurun_out/n22_tests_a.log
$T 60 python tools/gpu_batch_time.py > gpurun_out/n22_time.log 2>&1
$T 90 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "22-34-3 or 22-3-33 or 24-33-3 or 24-3-33" > gpurun_out/n22_tests_b.log 2>&1; echo "exit $?" >> gpurun_out/n22_tests_b.log
$T 150 python bench.py --workload heattrans22 --steps 20 --warmup 3 > gpurun_out/bench_heattrans22_batch3.json 2> gpurun_out/bench_heattrans22_batch3.err
tail -3 gpurun_out/n22_tests_a.log gpurun_out/n22_tests_b.log; cat gpurun_out/n22_time.log; cat gpurun_out/bench_heattrans22_batch3.json
