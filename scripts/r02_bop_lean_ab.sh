#!/bin/bash
# lean bond table (8 blocks/SM) A/B on C4, 1 GPU; tests first
set -u
OUT=gpurun_out/r02_lean
mkdir -p $OUT
python -m pytest tests/test_gpu_md.py tests/test_gpu_bop.py tests/test_gpu_dd.py -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest.log
for lean in 1 0; do
ATX_BOP_LEAN=$lean python bench.py --steps 20 --warmup 5 --blocks c4 --no-cpu > $OUT/bench_l$lean.json 2> $OUT/bench_l$lean.err
python - $lean <<'PY' | tee -a $OUT/summary.txt
import json, sys
d = json.loads(open('gpurun_out/r02_lean/bench_l%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
print('lean', sys.argv[1], 'C2 %.1f M' % (d['value'] / 1e6))
for k, b in d['c4'].items():
    print('   C4', k, '%.4f ms/step' % b['ms_per_step'], 'epot/atom', b['epot_per_atom'], {x: round(y, 4) for x, y in b['scopes_ms_per_step_rank0'].items() if y > 0.01}, b['roofline']['fp64']['frac'])
PY
done
