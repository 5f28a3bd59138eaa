import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from atomistica_b200 import native, structures as S
d = dict(np.load('tests/golden/aC.npz', allow_pickle=False))
a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True).repeat(3)
p = native.from_atoms(a); nl = native.Neighbors(50); pot = native.Rebo2(); pot.bind_to(p, nl)
for _ in range(3): e = pot.energy_and_forces(p, nl)[0]
print(len(a), e / len(a))
