#!/bin/bash
# final 1-GPU measurements of the round: GPU suite, both bench arms with the driver's flags, launch list,
# ncu summaries of the EAM and REBO2 kernels (CSV on the box, reports deleted)
set -u
OUT=gpurun_out/r02_final1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt; tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_n1.json 2> $OUT/bench_ref_n1.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "bench exit $?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --gpus 1 --steps 1000 --warmup 50 --blocks c2 --no-cpu > $OUT/bench_n1_1000.json 2> $OUT/bench_n1_1000.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for n in ('bench_ref_n1', 'bench_n1', 'bench_n1_1000'):
    try:
        d = json.loads(open('gpurun_out/r02_final1/%s.json' % n).read().strip().split('\n')[-1])
        print(n, 'value %.1f M  e2e %.1f M  ms/step %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d.get('ms_per_step')))
        if 'roofline' in d: print('  roofline frac %.3f avg %.4f ms whole %.3f fp64 %s' % (d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['whole_step_frac'], d['roofline'].get('fp64')), 'steady', d['steady_state']['value'] / 1e6)
        if d.get('c1'): print('  c1 %.3f ms %.1f M' % (d['c1']['ms_per_call'], d['c1']['value'] / 1e6), d['c1'].get('cpu_baseline', {}).get('value'))
        if d.get('c3'): print('  c3', d['c3']['device_ms'], '%.1f M' % (d['c3']['value'] / 1e6), d['c3']['roofline'].get('fp64'), d['c3'].get('cpu_baseline', {}).get('value'))
        for k, b in (d.get('c4') or {}).items(): print('  c4', k, '%.1f M %.4f ms/step reb %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b['roofline'].get('fp64'))
        if d.get('nl_sweep'):
            for r in d['nl_sweep']['rows']: print('  nl', r['system'], r['atoms'], r['cutoff'], '%.3f ms frac %.3f' % (r['ms'], r['hbm_frac']))
        print('  cpu', (d.get('cpu_baseline') or {}).get('value'), 'clocks', d.get('clocks'), 'errors', d.get('block_errors'))
    except Exception as e:
        print(n, 'failed', e)
PY
prof () {
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -c $cnt -k regex:"$rx" -o $OUT/$name -f "$@" > $OUT/ncu_$name.log 2>&1
  echo "  ncu $name exit $?" | tee -a $OUT/summary.txt
  python scripts/summarize_ncu.py full $OUT/$name.ncu-rep $OUT/r02_ncu_$name.csv > /dev/null 2>> $OUT/summary.txt
  rm -f $OUT/$name.ncu-rep
}
prof eam_final 'k_eam_' 4 python bench.py --steps 5 --warmup 3 --blocks c2 --no-cpu
cp profiles/traffic.json $OUT/traffic.json 2>/dev/null
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
prof rebo2_final 'k_rebo2_' 4 python /tmp/prof_rebo2.py
prof nl_cu_final 'k_pairs_coop|k_rows_to_csr|k_gather_sorted|k_cell_' 14 python scripts/r02_nl_probe.py
prof bop_final 'k_bop_center|k_bop_gather' 6 python bench.py --steps 5 --warmup 3 --blocks c4 --c4-kinds Tersoff --c4-cells 64 --c4-steps 10 --no-cpu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_raw.csv python bench.py --steps 20 --warmup 5 --blocks c2 --no-cpu > $OUT/ncu_launches.log 2>&1
python scripts/summarize_ncu.py launches $OUT/launches_raw.csv $OUT/r02_launches_final.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 20 --warmup 5 --blocks c2 --no-cpu" > /dev/null 2>> $OUT/summary.txt
rm -f $OUT/launches_raw.csv
du -sh $OUT | tee -a $OUT/summary.txt
