#!/bin/bash
# 1-GPU call: full GPU suite (new pair search), driver-flag bench with every block, ncu captures
# summarised to CSV ON THE BOX (the .ncu-rep files are deleted: gpurun_out is capped at 64 MiB).
set -u
OUT=gpurun_out/r02_gpu1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== 1. GPU suite" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt; tail -8 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== 2. bench, driver flags" | tee -a $OUT/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "exit $?" | tee -a $OUT/summary.txt; tail -3 $OUT/bench_n1.err | tee -a $OUT/summary.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_n1.json 2> $OUT/bench_ref_n1.err
echo "ref exit $?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --gpus 1 --steps 1000 --warmup 50 --blocks c2 --no-cpu > $OUT/bench_n1_1000.json 2> $OUT/bench_n1_1000.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for n in ('bench_n1', 'bench_n1_1000', 'bench_ref_n1'):
    try:
        d = json.loads(open('gpurun_out/r02_gpu1/%s.json' % n).read().strip().split('\n')[-1])
        print(n, 'value %.1f M  e2e %.1f M  ms/step %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d.get('ms_per_step')))
        if 'roofline' in d: print('  roofline', {k: d['roofline'][k] for k in ('frac', 'avg_launch_ms', 'whole_step_frac')}, 'steady', d.get('steady_state'))
        for k in ('c1', 'c3'):
            if d.get(k): print(' ', k, {q: d[k][q] for q in d[k] if q not in ('workload',)})
        for k, b in (d.get('c4') or {}).items(): print('  c4', k, {q: b[q] for q in b if q not in ('workload',)})
        if d.get('nl_sweep'):
            for r in d['nl_sweep']['rows']: print('  nl', r)
            print('  nl cpu', d['nl_sweep'].get('cpu_baseline'))
        print('  cpu', d.get('cpu_baseline')); print('  clocks', d.get('clocks'), 'errors', d.get('block_errors'))
    except Exception as e:
        print(n, 'failed', e)
PY
echo "== 3. REBO2 per-bond A/B (C3)" | tee -a $OUT/summary.txt
for v in 0 1 2 3; do
  ATX_REBO2_PERBOND=$v timeout 300 python benchmarks/run_configs.py C3 --out $OUT/c3_perbond$v.json > $OUT/c3_perbond$v.log 2>&1
  python -c "import json;d=json.load(open('$OUT/c3_perbond$v.json'))['C3'];print('  per_bond=$v', d['device_ms'], d['energy_per_atom'])" | tee -a $OUT/summary.txt
done
echo "== 4. ncu" | tee -a $OUT/summary.txt
cat > /tmp/prof_nl.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, bench
from atomistica_b200 import native
which = sys.argv[1]
if which == 'cu':
    pos, cell = bench.fcc_positions(3.615, 100); Z, sig, rc, avgn = 29, 0.35, 5.50679, 200
else:
    pos, cell = bench.diamond_positions(5.432, 80); Z, sig, rc, avgn = 14, 0.05, 3.0, 40
pos += np.random.RandomState(1).normal(scale=sig, size=pos.shape)
p = native.from_arrays(np.full(len(pos), Z, dtype=np.int32), pos, cell, True)
nl = native.Neighbors(avgn); nl.request_interaction_range(rc)
for _ in range(3): nl.rebuild(p)
print(which, len(pos), nl.info())
PY
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
prof () {  # name, kernel regex, count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -c $cnt -k regex:"$rx" -o $OUT/$name -f "$@" > $OUT/ncu_$name.log 2>&1
  echo "  ncu $name exit $?" | tee -a $OUT/summary.txt
  python scripts/summarize_ncu.py full $OUT/$name.ncu-rep $OUT/r02_ncu_$name.csv > /dev/null 2>> $OUT/summary.txt
  rm -f $OUT/$name.ncu-rep
}
prof eam 'k_eam_' 6 python bench.py --steps 10 --warmup 3 --blocks c2 --no-cpu
prof nl_cu 'k_pairs|k_cell|k_gather|k_rows|k_count' 14 python /tmp/prof_nl.py cu
prof nl_si 'k_pairs|k_cell|k_gather|k_rows|k_count' 14 python /tmp/prof_nl.py si
prof bop_tersoff 'k_bop_center|k_bop_gather' 4 python scripts/run_bop_md.py Tersoff 64 4
prof bop_kumagai 'k_bop_center|k_bop_gather' 4 python scripts/run_bop_md.py Kumagai 64 4
prof rebo2 'k_rebo2_' 6 python /tmp/prof_rebo2.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_raw.csv python bench.py --steps 20 --warmup 5 --blocks c2 --no-cpu > $OUT/ncu_launches.log 2>&1
python scripts/summarize_ncu.py launches $OUT/launches_raw.csv $OUT/r02_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 20 --warmup 5 --blocks c2 --no-cpu" > /dev/null 2>> $OUT/summary.txt
rm -f $OUT/launches_raw.csv
du -sh $OUT | tee -a $OUT/summary.txt
