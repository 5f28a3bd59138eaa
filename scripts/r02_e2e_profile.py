"""e2e path of bench.py (calculator API, 256000 Cu atoms) under cProfile, with and without the in-place
positions buffer; prints ms per call and the top of the profile.  Test/measurement helper."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from atomistica_b200 import TabulatedAlloyEAM, md, structures as S  # noqa: E402


def run(alias, profile):
    setfl = bench.load_setfl() if hasattr(bench, 'load_setfl') else dict(np.load(os.path.join(bench.ROOT, 'tests', 'golden', 'cu_mishin1_setfl.npz'), allow_pickle=False))
    calc = TabulatedAlloyEAM(setfl=setfl, device=0, verlet_shell=bench.SKIN, alias_positions=alias)
    pos0, cell0 = bench.fcc_positions(bench.A0, bench.NCELL)
    nat = len(pos0)
    a2 = S.Atoms(['Cu'] * nat, pos0, cell0, True)
    r = a2.positions
    m = np.full(nat, bench.MASS_CU)
    v = md.maxwell_boltzmann(m, bench.TEMP, seed=12345)
    f = calc.get_forces(a2)
    t_api = 0.0
    pr = cProfile.Profile() if profile else None
    n = 30
    for k in range(3 + n):
        v += 0.5 * f / bench.MASS_CU * md.ACCEL_CONV * bench.DT
        r += v * bench.DT
        t0 = time.perf_counter()
        if pr and k >= 3:
            pr.enable()
        f = calc.get_forces(a2)
        if pr and k >= 3:
            pr.disable()
        dt = time.perf_counter() - t0
        v += 0.5 * f / bench.MASS_CU * md.ACCEL_CONV * bench.DT
        if k >= 3:
            t_api += dt
    print('alias=%d profile=%d: %.3f ms per call = %.1f M atom-steps/s' % (alias, profile, 1e3 * t_api / n, nat * n / t_api / 1e6))
    if pr:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14)
        print(s.getvalue())


for alias in (0, 1):
    run(alias, False)
for alias in (1,):
    run(alias, True)
run(0, False)
run(1, False)
