#!/usr/bin/env bash
# 8-GPU scaling session (one box): N = 1, 2, 4, 8 for the three workloads
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # n workload extra...
  n=$1; wl=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --workload $wl --no-cpu "$@" 2>&1 | tail -1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload $wl --no-cpu "$@" 2>&1 | grep '^{' | tail -1
  fi
}
: > gpurun_out/scale_r01.jsonl
for n in 1 2 4 8; do run $n 0d_n32 --steps 50 --warmup 5 --weights synthetic >> gpurun_out/scale_r01.jsonl; done
for n in 1 2 4 8; do run $n shock1p2 --steps 10 --warmup 3 >> gpurun_out/scale_r01.jsonl; done
for n in 1 2 4 8; do run $n heattrans --steps 10 --warmup 3 >> gpurun_out/scale_r01.jsonl; done
for n in 8; do SBTE_CELLS_PER_GPU=80 run $n shock1p2 --steps 20 --warmup 3 >> gpurun_out/scale_r01.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/scale_r01.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:200]); continue
    print(d['config']['workload'][:40], 'N=',d['n_gpus'], 'value=%.1f'%d['value'], d['unit'], 'ms/step=%.3f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], 'e2e=%.1f'%d['e2e']['value'], 'clk', (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'))
PY
