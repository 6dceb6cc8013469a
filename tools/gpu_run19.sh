#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_0d_cluster.csv python bench.py --steps 3 --warmup 3 --no-cpu --weights synthetic > gpurun_out/b.log 2>&1
grep -c qhat gpurun_out/launches_0d_cluster.csv
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_0d_cluster.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:40]:
    print(r[ki][:50], r[vi], r[ui])
PY
cat > /tmp/n24.py <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
for N in (24, 16):
    c = sb.Collisions(N, 5.0)
    c.synthetic_weights(1)
    f = initial.init_hom(c.v, 5.0, 0)
    d = c.array(c.n3).put(f); q = c.array(c.n3)
    for _ in range(5): sb._lib.check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, 0))
    c.sync(); t0 = time.perf_counter()
    for _ in range(200): sb._lib.check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, 0))
    c.sync(); print("N", N, "ComputeQ device ms", (time.perf_counter() - t0) / 200 * 1e3)
PY
timeout 300 python /tmp/n24.py
SBTE_NO_CLUSTER_FFT=1 timeout 300 python /tmp/n24.py
