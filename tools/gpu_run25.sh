#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4 | cut -c1-300
echo "=== shock"; timeout 600 python bench.py --workload shock1p2 --steps 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_shock_b.csv python bench.py --workload shock1p2 --steps 2 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_shock_b.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[2:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    u=r[ui]
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    agg[r[ki][:60]][0]+=1; agg[r[ki][:60]][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:62s} n={v[0]:4d} avg={v[1]/v[0]:8.1f}us share={v[1]/tot:.3f}")
PY
