#!/bin/bash
set -u
OUT=gpurun_out/r02_gpu1d
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== 1. GPU suite" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt; tail -12 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== 2. bench N=1: c2, c3, c4" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --blocks c3,c4 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open('gpurun_out/r02_gpu1d/bench_n1.json').read().strip().split('\n')[-1])
print('C2 value %.1f M steady %.1f M e2e %.1f M' % (d['value'] / 1e6, d['steady_state']['value'] / 1e6, d['e2e']['value'] / 1e6))
print('  c3', d['c3']['device_ms'], d['c3']['roofline'])
for k, b in (d.get('c4') or {}).items():
    print('   C4', k, '%.1f M, ms/step %.4f rebuilds %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b['scopes_ms_per_step_rank0'], b['roofline'].get('fp64'))
print(d.get('block_errors'))
PY
