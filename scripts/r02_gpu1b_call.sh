#!/bin/bash
# 1-GPU call: seam-2 module tests, neighbour sweep after the hit-path change, REBO2 occupancy A/B,
# ncu with explicit FP64 instruction counts (flops per atom for the roofline objects).
set -u
OUT=gpurun_out/r02_gpu1b
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== 1. seam 2 + neighbour + rebo2 + md tests" | tee $OUT/summary.txt
timeout 900 python -m pytest tests/test_seam2.py tests/test_gpu_neighbors.py tests/test_gpu_rebo2.py tests/test_gpu_md.py tests/test_gpu_external_list.py -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt; tail -25 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== 2. neighbour sweep" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --blocks nl --no-cpu --nl-sizes 1e6,4e6 > $OUT/bench_nl.json 2> $OUT/bench_nl.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open('gpurun_out/r02_gpu1b/bench_nl.json').read().strip().split('\n')[-1])
print('value %.1f M e2e %.1f M' % (d['value'] / 1e6, d['e2e']['value'] / 1e6))
for r in d['nl_sweep']['rows']: print('  nl', r['system'], r['atoms'], r['cutoff'], '%.3f ms  frac %.3f' % (r['ms'], r['hbm_frac']))
PY
echo "== 3. REBO2 occupancy A/B (C3)" | tee -a $OUT/summary.txt
for v in 3 4 5; do
  ATX_REBO2_PERBOND=$v timeout 300 python benchmarks/run_configs.py C3 --out $OUT/c3_perbond$v.json > $OUT/c3_perbond$v.log 2>&1
  python -c "import json;d=json.load(open('$OUT/c3_perbond$v.json'))['C3'];print('  per_bond=$v', d['device_ms'], d['energy_per_atom'])" | tee -a $OUT/summary.txt
done
echo "== 4. ncu (FP64 instruction counts)" | tee -a $OUT/summary.txt
cat > /tmp/prof_rebo2.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from atomistica_b200 import native, structures as S
d = dict(np.load('tests/golden/aC.npz', allow_pickle=False))
a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True).repeat(3)
p = native.from_atoms(a); nl = native.Neighbors(50); pot = native.Rebo2(); pot.bind_to(p, nl)
for _ in range(3): e = pot.energy_and_forces(p, nl)[0]
print(len(a), e / len(a))
PY
M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_global_red.sum,dram__bytes_read.sum,dram__bytes_write.sum
prof () {
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --metrics $M --clock-control none -c $cnt -k regex:"$rx" --csv --log-file $OUT/r02_flops_$name.csv "$@" > $OUT/ncu_$name.log 2>&1
  echo "  ncu $name exit $?" | tee -a $OUT/summary.txt
}
prof bop_tersoff 'k_bop_center<' 2 python scripts/run_bop_md.py Tersoff 64 2
prof bop_kumagai 'k_bop_center<' 2 python scripts/run_bop_md.py Kumagai 64 2
prof rebo2 'k_rebo2_force' 2 python /tmp/prof_rebo2.py
prof eam 'k_eam_' 4 python bench.py --steps 5 --warmup 3 --blocks c2 --no-cpu
timeout 600 ncu --set full --clock-control none --import-source on -c 2 -k regex:'k_rebo2_force' -o $OUT/rebo2b -f python /tmp/prof_rebo2.py > $OUT/ncu_rebo2b.log 2>&1
python scripts/summarize_ncu.py full $OUT/rebo2b.ncu-rep $OUT/r02_ncu_rebo2_perbond.csv > /dev/null 2>> $OUT/summary.txt
ncu -i $OUT/rebo2b.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > $OUT/rebo2b_source.csv
rm -f $OUT/rebo2b.ncu-rep
du -sh $OUT | tee -a $OUT/summary.txt
