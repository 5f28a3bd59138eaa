#!/bin/bash
# 8-GPU diagnostic: C4 Tersoff only, no parity runs, per-scope times of rank 0, clocks of all GPUs
set -u
OUT=gpurun_out/r02_n8diag
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 250 > $OUT/smi.csv 2>/dev/null &
SMI=$!
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 --blocks c4 --c4-kinds Tersoff --no-parity --c4-steps 120 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "exit $?" | tee $OUT/summary.txt
kill $SMI
python - <<'PY' | tee -a $OUT/summary.txt
import json, collections
d = json.loads(open('gpurun_out/r02_n8diag/bench_n8.json').read().strip().split('\n')[-1])
print('C2 N=8 %.1f M  steady %.1f M' % (d['value'] / 1e6, d['steady_state']['value'] / 1e6), d['kernels_ms']['dd'])
for k, b in d['c4'].items():
    print('C4', k, '%.1f M ms/step %.4f rebuilds %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b['scopes_ms_per_step_rank0'], b['rebuild_host_ms_since_create'])
g = collections.defaultdict(list)
for l in open('gpurun_out/r02_n8diag/smi.csv'):
    t = [x.strip() for x in l.split(',')]
    try: g[t[0]].append((float(t[1]), float(t[2]), t[3]))
    except Exception: pass
for k in sorted(g):
    v = [x for x in g[k] if x[1] > 300]
    if v: print('gpu', k, 'busy samples', len(v), 'sm MHz min/median', min(x[0] for x in v), sorted(x[0] for x in v)[len(v) // 2], 'power max', max(x[1] for x in v), 'power cap', sum(1 for x in v if x[2].startswith('Active')))
PY
