#!/bin/bash
# (tools/ab/ is not tracked: build the old library with `git worktree add /tmp/wt 2d82f08 && (cd /tmp/wt && python -m spectralbte_b200.build)`
#  and copy /tmp/wt/spectralbte_b200/libsbte_b200.so to tools/ab/libsbte_precanon.so first.)
# Same-box A/B of the batched convolution: the current library (canonical summation order, fold into part 0 by
# red.global.add.f64) against the library before the canonical order (tools/ab/libsbte_precanon.so, commit 2d82f08).
mkdir -p gpurun_out
for cfg in "24 250" "22 250" "16 640" "16 80" "8 256"; do
  set -- $cfg
  timeout 90 python tools/gpu_batch_time.py $1 $2
  SBTE_LIB_PATH=$PWD/tools/ab/libsbte_precanon.so timeout 90 python tools/gpu_batch_time.py $1 $2 | sed "s/batched kernel/PRE-CANONICAL/"
done 2>&1 | tee gpurun_out/r02_canon_ab3.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "split_tiles or batched_computeq or two_rank or peer_memory or 1d_step or shock1p2 or line_ring or heattrans or golden" > gpurun_out/r02_red_tests.log 2>&1
tail -3 gpurun_out/r02_red_tests.log
