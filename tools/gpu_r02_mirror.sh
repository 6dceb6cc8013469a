#!/bin/bash
# First GPU call of the next round: the mirror-paired batched convolution (SBTE_MIRROR=1, csrc/qhat_mirror.cu) has only
# been checked on the CPU (tests/test_mirror_emulation_cpu.py).  1) the batched / 1D parity tests of the suite with the
# kernel switched on, 2) time against the default kernel at N=16 (640 and 80 cells), 3) the rolled N=24 line-ring
# variant (SBTE_ROLL=1) against the default at 250 cells.  Every step under its own timeout (a wrong barrier count
# would hang, not fail).
mkdir -p gpurun_out
SBTE_MIRROR=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "batched_computeq or symmetrised_stream or 1d_step or heat_transport_golden or shock1p2" > gpurun_out/mirror_tests.log 2>&1
echo "exit $?" >> gpurun_out/mirror_tests.log
for m in 0 1 3; do   # 3 = mirror kernel on the folded tensor (combined body on the foldable steps), only on the way to Q
  SBTE_MIRROR=$m timeout 40 python tools/gpu_n22_time.py 16 640 >> gpurun_out/mirror_time.log 2>&1
  SBTE_MIRROR=$m timeout 40 python tools/gpu_n22_time.py 16 80 >> gpurun_out/mirror_time.log 2>&1
done
SBTE_MIRROR=3 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "batched_computeq or symmetrised_stream or 1d_step or heat_transport_golden or shock1p2" > gpurun_out/mirror_fold_tests.log 2>&1
echo "exit $?" >> gpurun_out/mirror_fold_tests.log
for r in 0 1; do
  SBTE_ROLL=$r timeout 40 python tools/gpu_n22_time.py 24 250 >> gpurun_out/roll_time.log 2>&1
  SBTE_ROLL=$r timeout 40 python tools/gpu_n22_time.py 22 250 >> gpurun_out/roll_time.log 2>&1
done
SBTE_ROLL=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "line_ring or 24-33-3 or 24-3-33 or 22-34-3 or 22-3-33" > gpurun_out/roll_tests.log 2>&1; echo "exit $?" >> gpurun_out/roll_tests.log
# the line-ring mirror kernels (N = 20, 22, 24) need SBTE_MIRROR=2
SBTE_MIRROR=2 timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "line_ring or 24-33-3 or 24-3-33 or 22-34-3 or 22-3-33" > gpurun_out/mirror_ring_tests.log 2>&1
echo "exit $?" >> gpurun_out/mirror_ring_tests.log
for n in 24 22 20; do SBTE_MIRROR=2 timeout 40 python tools/gpu_n22_time.py $n 250 >> gpurun_out/mirror_time.log 2>&1; done
tail -n 3 gpurun_out/mirror_tests.log gpurun_out/mirror_fold_tests.log gpurun_out/mirror_ring_tests.log; cat gpurun_out/mirror_time.log gpurun_out/roll_time.log
# sanitizer pass over the new kernels (small cases): memcheck, then racecheck (shared-memory hazards, barrier misuse)
SBTE_MIRROR=2 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "8-5-3 or 8-37-3 or 16-33-3" > gpurun_out/mirror_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/mirror_memcheck.log
SBTE_MIRROR=2 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "8-5-3" > gpurun_out/mirror_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/mirror_racecheck.log
tail -n 4 gpurun_out/mirror_memcheck.log gpurun_out/mirror_racecheck.log
# one full ncu capture of each new kernel and of the kernels they would replace (FP64 pipe, stall reasons, LSU wavefronts,
# instruction-fetch stalls): read offline with tools/ncu_summary.py
for cfg in "1 16 640 mirror_n16" "0 16 640 batch2_n16" "2 24 250 mirror_ring_n24" "0 24 250 batch3_n24"; do
  set -- $cfg
  SBTE_MIRROR=$1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:"qhat_(batch2|batch3|mirror)" -s 2 -c 1 \
    -o gpurun_out/r02_k2_$4 -f python tools/gpu_n22_time.py $2 $3 > gpurun_out/ncu_$4.log 2>&1
done
SBTE_ROLL=1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:qhat_batch3 -s 2 -c 1 \
  -o gpurun_out/r02_k2_batch3_n24_rolled -f python tools/gpu_n22_time.py 24 250 > gpurun_out/ncu_batch3_n24_rolled.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 0D on half of the zeta rows (SBTE_HALF0D=1, csrc/qhat_half.cu): parity of ComputeQ at N = 16 / 32 against the oracle, then
# the headline bench with and without it
SBTE_HALF0D=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider \
  -k "computeq_and_maxpreserve or n32 or conservation_to_1e13 or bkw16" > gpurun_out/half0d_tests.log 2>&1; echo "exit $?" >> gpurun_out/half0d_tests.log
for h in 0 1; do SBTE_HALF0D=$h timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_half0d_$h.json 2> gpurun_out/bench_half0d_$h.err; done
tail -n 3 gpurun_out/half0d_tests.log; cut -c1-400 gpurun_out/bench_half0d_0.json gpurun_out/bench_half0d_1.json
