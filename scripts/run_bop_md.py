"""Small driver for profiling: Tersoff (or Kumagai) Si NVE steps on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from atomistica_b200 import md, native, structures as S
kind = sys.argv[1] if len(sys.argv) > 1 else 'Tersoff'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
a = S.diamond('Si', 5.432 if kind == 'Tersoff' else 5.429, (n, n, n))
a.rattle(0.05, seed=12345)
m = np.full(len(a), 28.0855)
v0 = md.maxwell_boltzmann(m, 300.0, seed=12345)
p = native.from_atoms(a)
nl = native.Neighbors(20)
pot = getattr(native, kind)()
drv = md.VelocityVerlet(pot, p, nl, m, v0, dt=1.0, verlet_shell=0.4)
drv.run(3)
e = drv.run(steps)
st = drv.stats()
print(kind, len(a), 'atoms', st['last_run_ms'] / steps, 'ms/step', len(a) * steps / st['last_run_ms'] / 1e3, 'M atom-steps/s', st, e[0] / len(a))
