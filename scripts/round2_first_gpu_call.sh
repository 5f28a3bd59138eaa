#!/bin/bash
# First GPU call of the next round (one gpurun, 1 GPU, about 20 minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_gpu_call.sh'
# 1. runs what round 1 could only check on the CPU (fenced with ATX_RUN_UNVERIFIED): the Rebo2Scr
#    kernels and the consecutive lane mapping of the EAM kernels;
# 2. A/B of the EAM lane mapping and lanes per atom on the bench configuration (tells the two L1
#    wavefront rules of DESIGN.md section 8.1 apart);
# 3. first ncu captures of the kernels that were never profiled: REBO2, Rebo2Scr, screened BOP.
# Second call, 2 GPUs (REBO2 under domain decomposition, also fenced):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'ATX_RUN_UNVERIFIED=1 python -m pytest tests/test_gpu_dd.py -m gpu -q;
#     python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
#       benchmarks/run_dd.py --kind Rebo2 --cells 5 --steps 50 --skin 0.3'     # C3 under decomposition
# Everything lands in gpurun_out/r02_first/; nothing here is a bench value (ncu runs are profiles).
set -u
OUT=gpurun_out/r02_first
mkdir -p $OUT
export PYTHONUNBUFFERED=1

echo "== 1. full GPU suite incl. fenced tests" | tee $OUT/summary.txt
ATX_RUN_UNVERIFIED=1 timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_unverified.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_unverified.log | tee -a $OUT/summary.txt

echo "== 2. EAM lane mapping A/B (bench.py, 300 steps)" | tee -a $OUT/summary.txt
for cfg in "0 4 2" "1 4 2" "0 8 2" "1 8 2" "1 16 1"; do
  set -- $cfg
  ATX_EAM_MAP=$1 ATX_EAM_LANES=$2 ATX_EAM_UNROLL=$3 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu \
    > $OUT/bench_map$1_l$2_u$3.json 2> $OUT/bench_map$1_l$2_u$3.err
  python - "$OUT/bench_map$1_l$2_u$3.json" "$cfg" <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    n = max(d['roofline']['launches'], 1)
    print('  map/lanes/unroll %s: %.1f M atom-steps/s, force %.1f us, density %.1f us' % (
        sys.argv[2], d['value'] / 1e6, d['roofline']['avg_launch_ms'] * 1e3, d['kernels_ms']['eam_density'] / n * 1e3))
except Exception as e:
    print('  %s: failed (%s)' % (sys.argv[2], e))
PY
done

echo "== 2b. REBO2 C3: thread per atom (0) vs thread per bond at 4 / 6 / 8 blocks per SM (1 / 2 / 3)" | tee -a $OUT/summary.txt
for v in 0 1 2 3; do
  ATX_REBO2_PERBOND=$v timeout 600 python benchmarks/run_configs.py C3 --out $OUT/c3_perbond$v.json > $OUT/c3_perbond$v.log 2>&1
  grep -o '"device_ms": {[^}]*}' $OUT/c3_perbond$v.json | head -1 | sed "s/^/  per_bond=$v /" | tee -a $OUT/summary.txt
done

echo "== 3. ncu: REBO2 (C3, small replica), Rebo2Scr, screened BOP" | tee -a $OUT/summary.txt
cat > /tmp/r02_prof.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from atomistica_b200 import native, structures as S, parameters as P
d = dict(np.load('tests/golden/aC.npz', allow_pickle=False))
a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True).repeat(3)
which = sys.argv[1]
p = native.from_atoms(a)
if which == 'rebo2':
    nl, pot = native.Neighbors(50), native.Rebo2()
elif which == 'rebo2scr':
    nl, pot = native.Neighbors(1000), native.Rebo2Scr()
else:
    nl, pot = native.Neighbors(1000), native.TersoffScr()
pot.bind_to(p, nl)
for _ in range(3):
    e = pot.energy_and_forces(p, nl)[0]
print(which, len(a), 'atoms, epot/atom', e / len(a))
PY
for w in rebo2 rebo2scr tersoffscr; do
  timeout 600 ncu --set full --clock-control none --import-source on -c 12 \
    -k regex:'k_rebo2_force|k_rebo2_bonds|k_rbs_|k_bopscr_' -o $OUT/r02_$w -f python /tmp/r02_prof.py $w \
    > $OUT/ncu_$w.log 2>&1
  echo "  ncu $w exit $?" | tee -a $OUT/summary.txt
  tail -2 $OUT/ncu_$w.log | tee -a $OUT/summary.txt
done
# neighbour build (never profiled with --set full): two builds of the bench system
timeout 600 ncu --set full --clock-control none --import-source on -c 6 -k regex:'k_pairs|k_rows_to_csr' \
  -o $OUT/r02_nl -f python bench.py --steps 70 --warmup 3 --no-cpu > $OUT/ncu_nl.log 2>&1
echo "  ncu nl exit $?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
# final k_bop_center instantiation (never captured): Tersoff and Kumagai Si, 2.1 M atoms
for k in Tersoff Kumagai; do
  timeout 600 ncu --set full --clock-control none --import-source on -c 4 -k regex:'k_bop_center|k_bop_gather' \
    -o $OUT/r02_bop_$k -f python scripts/run_bop_md.py $k 64 4 > $OUT/ncu_bop_$k.log 2>&1
  echo "  ncu bop $k exit $?" | tee -a $OUT/summary.txt
done
python scripts/run_bop_md.py Tersoff 64 50 | tee -a $OUT/summary.txt
python scripts/run_bop_md.py Kumagai 64 50 | tee -a $OUT/summary.txt
nvidia-smi -L | tee -a $OUT/summary.txt
which gfortran | tee -a $OUT/summary.txt
