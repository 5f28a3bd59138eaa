#!/bin/bash
# 2-GPU diagnostic at the per-rank size of the 8-GPU C4 run (80^3 cells = 4.1 M atoms, 2.05 M per rank):
# exact rebuild phases (ATX_DD_PROFILE=1 synchronises around every phase) and per-scope step times
set -u
OUT=gpurun_out/r02_n2diag
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for prof in 1 0; do
ATX_DD_PROFILE=$prof timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --blocks c4 --c4-kinds Tersoff --no-parity --c4-steps 120 --c4-cells 80 > $OUT/bench_p$prof.json 2> $OUT/bench_p$prof.err
echo "exit $?" | tee -a $OUT/summary.txt
python - $prof <<'PY' | tee -a $OUT/summary.txt
import json, sys
d = json.loads(open('gpurun_out/r02_n2diag/bench_p%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
for k, b in d['c4'].items():
    print('profile', sys.argv[1], 'C4', k, '%.1f M ms/step %.4f rebuilds %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b.get('scopes_ms_per_step_rank0'), b['rebuild_host_ms_since_create'], b.get('priming_steps_untimed'))
PY
done
