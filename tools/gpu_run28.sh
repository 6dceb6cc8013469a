#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== mp32 script"; timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
echo "=== bench synthetic"; timeout 600 python bench.py --no-cpu --weights synthetic 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('maxpreserve',{}).get('ms_per_call'), d['clocks'])"
echo "=== bench generated"; timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('maxpreserve',{}).get('ms_per_call'), d['clocks'])"
