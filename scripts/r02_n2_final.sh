#!/bin/bash
# the driver's command line at N=2 with the final code (C2 + parity, C4 Tersoff / Kumagai + parity)
set -u
OUT=gpurun_out/r02_n2final
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_dd.py tests/test_gpu_multidevice.py -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "bench n2 exit $?" | tee $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open('gpurun_out/r02_n2final/bench_n2.json').read().strip().split('\n')[-1])
print('C2 N=2 %.1f M %.4f ms/step steady %.1f M e2e %.1f M parity %s' % (d['value'] / 1e6, d['ms_per_step'], d['steady_state']['value'] / 1e6, d['e2e']['value'] / 1e6, d['parity']['ok']))
for k, b in d['c4'].items():
    print('C4', k, '%.1f M %.4f ms/step rebuilds %d parity %s' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds'], b['parity']['ok']), b.get('scopes_ms_per_step_rank0'))
print(d.get('block_errors'), d['clocks'])
PY
