#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -6 | cut -c1-300
echo "=== mp32 cluster"; timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
echo "=== mp32 passes"; SBTE_NO_CLUSTER_FFT=1 timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
for env in "" "SBTE_NO_CLUSTER_FFT=1"; do
echo "=== bench default $env"; env $env timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('maxpreserve'))"
echo "=== bkw16 $env"; env $env timeout 600 python bench.py --workload bkw16 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "=== heattrans $env"; env $env timeout 600 python bench.py --workload heattrans --steps 10 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
