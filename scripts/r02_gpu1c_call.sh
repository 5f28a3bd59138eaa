#!/bin/bash
set -u
OUT=gpurun_out/r02_gpu1c
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== 1. seam 2 + calculator tests" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_seam2.py tests/test_gpu_calculator.py -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt; tail -25 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== 2. C4 at N=1 (list cutoff 3.7)" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --blocks c4 --no-cpu > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open('gpurun_out/r02_gpu1c/bench_c4_n1.json').read().strip().split('\n')[-1])
print('C2 value %.1f M e2e %.1f M' % (d['value'] / 1e6, d['e2e']['value'] / 1e6), d['roofline'].get('fp64'))
for k, b in (d.get('c4') or {}).items():
    print('   C4', k, '%.1f M, ms/step %.4f rebuilds %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b['scopes_ms_per_step_rank0'], b.get('roofline'))
print(d.get('block_errors'))
PY
echo "== 3. ncu FP64 instruction counts, BOP" | tee -a $OUT/summary.txt
M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for k in Tersoff Kumagai; do
  timeout 600 ncu --metrics $M --clock-control none -c 2 -k regex:'^k_bop_center$' --csv --log-file $OUT/r02_flops_bop_$k.csv python scripts/run_bop_md.py $k 64 2 > $OUT/ncu_bop_$k.log 2>&1
  echo "  ncu bop $k exit $?" | tee -a $OUT/summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -c 2 -k regex:'k_pairs_f32' -o $OUT/nlb -f python - <<'PY' > $OUT/ncu_nlb.log 2>&1
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, bench
from atomistica_b200 import native
pos, cell = bench.fcc_positions(3.615, 100)
pos += np.random.RandomState(1).normal(scale=0.35, size=pos.shape)
p = native.from_arrays(np.full(len(pos), 29, dtype=np.int32), pos, cell, True)
nl = native.Neighbors(200); nl.request_interaction_range(5.50679)
for _ in range(3): nl.rebuild(p)
PY
python scripts/summarize_ncu.py full $OUT/nlb.ncu-rep $OUT/r02_ncu_nl_cu_v2.csv > /dev/null 2>> $OUT/summary.txt
rm -f $OUT/nlb.ncu-rep
du -sh $OUT | tee -a $OUT/summary.txt
