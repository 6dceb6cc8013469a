#!/bin/bash
# Round 2, final build: the full parity suite and the default bench command on one GPU.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_gpu_tests_1gpu_final.log 2>&1
tail -3 gpurun_out/r02_gpu_tests_1gpu_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_last_1gpu.json 2> gpurun_out/r02_bench_last_1gpu.err
echo "bench rc $?"; tail -c 300 gpurun_out/r02_bench_last_1gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_last_1gpu.json') if l.startswith('{')][-1])
print('0D value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])
for k,v in [(k,v) for k,v in d.get("oned",{}).items() if isinstance(v,dict)]:
    print(k, round(v['value']), 'ms/step', round(v['ms_per_step'],4), 'kernel', round(v.get('kernel_ms',0),4), 'rest', round(v.get('non_kernel_ms',0),4))
print('dropin', json.dumps(d.get('dropin'))[:600])
PY
