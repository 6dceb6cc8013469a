#!/bin/bash
# Short GPU call: row-split batched convolution (SBTE_ROW_SPLIT=1) against the default at N=16 -- time, bitwise
# comparison, and the N=16 oracle-parity cases of the suite with the row split on.
mkdir -p gpurun_out
timeout 40 python tools/gpu_rowsplit_ab.py > gpurun_out/rowsplit_ab.log 2>&1
if grep -q "bitwise equal True" gpurun_out/rowsplit_ab.log; then
  SBTE_ROW_SPLIT=1 timeout 25 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "16-33-3 or 16-3-40" > gpurun_out/rowsplit_tests.log 2>&1; echo "exit $?" >> gpurun_out/rowsplit_tests.log
fi
cat gpurun_out/rowsplit_ab.log; tail -n 3 gpurun_out/rowsplit_tests.log
