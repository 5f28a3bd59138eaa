"""per-call times of the calculator path (verlet_shell 0.5) over 120 MD-like calls: shows what a rebuild inside
the Verlet-shell mode costs.  Measurement helper."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from atomistica_b200 import TabulatedAlloyEAM, md, structures as S  # noqa: E402

setfl = bench.load_setfl()
calc = TabulatedAlloyEAM(setfl=setfl, device=0, verlet_shell=float(sys.argv[1]) if len(sys.argv) > 1 else 0.5)
pos0, cell0 = bench.fcc_positions(bench.A0, bench.NCELL)
nat = len(pos0)
a2 = S.Atoms(['Cu'] * nat, pos0, cell0, True)
r = a2.positions
v = md.maxwell_boltzmann(np.full(nat, bench.MASS_CU), bench.TEMP, seed=12345)
f = calc.get_forces(a2)
ts = []
for k in range(120):
    v += 0.5 * f / bench.MASS_CU * md.ACCEL_CONV * bench.DT
    r += v * bench.DT
    t0 = time.perf_counter()
    f = calc.get_forces(a2)
    ts.append(time.perf_counter() - t0)
    v += 0.5 * f / bench.MASS_CU * md.ACCEL_CONV * bench.DT
ts = np.array(ts) * 1e3
print('median %.3f ms' % np.median(ts), 'slow calls:', [(i, round(t, 2)) for i, t in enumerate(ts) if t > 1.5 * np.median(ts)])
print('builds', calc.nl.counters())
