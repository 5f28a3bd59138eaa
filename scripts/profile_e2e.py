"""Where does the time of one calculator call go? (C2 system, host buffers in/out)"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from atomistica_b200 import TabulatedAlloyEAM, _lib as L
import ctypes as C
setfl = bench.load_setfl()
a, m, v0 = bench.build_system()
calc = TabulatedAlloyEAM(setfl=setfl, verlet_shell=0.5)
a.calc = calc
calc.get_forces(a)
rng = np.random.RandomState(0)
def step():
    a.positions += rng.normal(scale=1e-3, size=a.positions.shape)
    calc.calculate(a)
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(10): step()
print('ms per call incl. rng', (time.perf_counter() - t0) / 10 * 1e3)
ctx = L.context(0)
L.check(L.lib().atx_profile_enable(ctx, 1))
pr = cProfile.Profile()
pr.enable()
for _ in range(10): step()
pr.disable()
L.check(L.lib().atx_profile_enable(ctx, 0))
for n in ('eam_density', 'eam_force', 'nl_pairs_count', 'nl_pairs_fill'):
    tot, cnt = C.c_double(0), C.c_longlong(0)
    L.lib().atx_profile_read(ctx, n.encode(), C.byref(tot), C.byref(cnt))
    print(n, tot.value / max(cnt.value, 1), 'ms')
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
