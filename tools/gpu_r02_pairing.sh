#!/bin/bash
# Round 2: transposed pairing of the 0D stream (half of the zeta columns): parity, memcheck on N=16, headline bench, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "transposed_pairing or n32 or computeq_and_maxpreserve or bkw16 or symmetrised_stream or conservation_to_1e13 or fft3d_matches_oracle" > gpurun_out/r02_pairing_tests.log 2>&1
tail -4 gpurun_out/r02_pairing_tests.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "transposed_pairing and 16-5.0" > gpurun_out/r02_pairing_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r02_pairing_memcheck.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-oned --no-dropin > gpurun_out/r02_bench_pairing.json 2> gpurun_out/r02_bench_pairing.err; echo "bench rc $?"; tail -c 300 gpurun_out/r02_bench_pairing.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_pairing.json'))
r=d['roofline'];print('value',d['value'],'e2e',d['e2e']['value'],'kernel_ms',r['kernel_ms'],'achieved',r['achieved'],'frac',r['frac'],'bytes',r['algorithmic_bytes_per_launch'])
print('sym_kernel',r['sym_kernel']);print('plain',r['plain_kernel']);print('xy',r['xy_pairing']);print('sustained',{k:(v['ms_per_call'],v['clocks']['sm_mhz']) for k,v in d.get('sustained',{}).items()})"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qhat_stream_kernel -s 5 -c 1 -o gpurun_out/r02_k2_stream_pairing_n32 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-dropin --no-sustained --no-oned > gpurun_out/ncu_pairing.log 2>&1
ls -la gpurun_out/r02_k2_stream_pairing_n32.ncu-rep
