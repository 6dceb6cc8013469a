#!/bin/bash
# Round 2: compute-sanitizer passes over the kernels that are new or changed this round (small cases): memcheck on the
# batched kernels with split tiles, the fused halo stencils (two slabs on one device), the paired stream kernels and the
# chained transform; racecheck on one batched case and one peer-halo case (TMA writes are reported by racecheck as
# hazards against mbarrier-ordered reads: those are listed, not errors of the protocol).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "split_tiles and 44 or peer_memory or two_rank and 16-8 or 1d_step or transposed_pairing and 16-5.0 or batched_computeq and 8-37 or heat_transport_golden or bkw8_golden" \
  > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/r02_memcheck.log
tail -4 gpurun_out/r02_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "peer_memory and 2-6 or 1d_step and 2-6" > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/r02_racecheck.log
grep -c "Race reported" gpurun_out/r02_racecheck.log; grep "Race reported" gpurun_out/r02_racecheck.log | sed 's/.*in \(.*\)$/\1/' | sort | uniq -c | head; tail -3 gpurun_out/r02_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "split_tiles and 12 or transposed_pairing and 16-5.0 or peer_memory and 2-6" > gpurun_out/r02_synccheck.log 2>&1; echo "synccheck exit $?" | tee -a gpurun_out/r02_synccheck.log
tail -3 gpurun_out/r02_synccheck.log
