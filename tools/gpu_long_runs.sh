#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --workload shock1p2 --steps 200 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r01_shock1p2_200steps.json
timeout 300 python bench.py --workload heattrans --steps 500 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r01_heattrans_500steps.json
python - <<'PY'
import json
for n in ("shock1p2_200steps","heattrans_500steps"):
    d=json.load(open("gpurun_out/bench_r01_%s.json"%n)); print(n, d["value"], d["ms_per_step"], d["clocks"], d["e2e"]["checksum"])
PY
