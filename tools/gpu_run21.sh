#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4 | cut -c1-300
echo "=== mp32"; timeout 300 python tools/gpu_mp32.py 2>&1 | tail -1
echo "=== n24/n16"; timeout 300 python tools/gpu_n24b.py
echo "=== bench default"; timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('maxpreserve'))"
echo "=== bkw16"; timeout 600 python bench.py --workload bkw16 --steps 50 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "=== heattrans"; timeout 600 python bench.py --workload heattrans --steps 10 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_mp32.csv python tools/gpu_mp32.py > gpurun_out/b.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_mp32.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[3:16]:
    print(r[ki][:50], r[vi], r[ui])
PY
