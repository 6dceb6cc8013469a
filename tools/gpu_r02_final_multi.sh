#!/bin/bash
# Round 2, final build: cross-GPU tests (when 2 GPUs) and the default bench command at N GPUs (N = $1), as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = 2 ] && [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider -k "two_gpus or peer or two_rank or invariance" > gpurun_out/r02_gpu_tests_2gpu_final.log 2>&1
  tail -3 gpurun_out/r02_gpu_tests_2gpu_final.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_last_${N}gpu.json 2> gpurun_out/r02_bench_last_${N}gpu.err
echo "bench rc $?"; tail -c 400 gpurun_out/r02_bench_last_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_last_${N}gpu.json') if l.startswith('{')][-1])
print("0D value", d["value"], "halo_parity", d.get("halo_parity"))
for k,v in [(k,v) for k,v in d.get("oned",{}).items() if isinstance(v,dict)]:
    print(k, round(v['value']), 'ms/step', round(v['ms_per_step'],4), 'kernel', round(v.get('kernel_ms',0),4), 'rest', round(v.get('non_kernel_ms',0),4), v.get('halo'))
PY
