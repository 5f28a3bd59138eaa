"""The reference's tests/test_forces_and_virial.py on the GPU calculators.

Same table (the rows of the hot path: pair styles, Brenner incl. the Fe-C set with masks, Kumagai,
Tersoff, the screened classes, Rebo2 / Rebo2Scr, both EAM classes), same procedure (:204-360): every
material is shifted by 0.1, checked in equilibrium and after a rattle, with the random masks where
the reference uses them; forces and stress against the finite-difference helpers of
`atomistica_b200.tests` (mirror of `atomistica.tests`), dx = 1e-6, tolerance 1e-2 on the reference's
error measures.

"""
import os

import numpy as np
import pytest

import atomistica_b200 as ab
from atomistica_b200 import structures as S
from atomistica_b200.tests import test_forces as forces, test_virial as virial
from conftest import load_npz

pytestmark = [pytest.mark.gpu]

sx = 2
dx = 1e-6
tol = 1e-2


def random_solid(els, density, seed=0):
    """tests/test_forces_and_virial.py:50-64: random positions at a mass density (g/cm^3)"""
    from atomistica_b200.elements import atomic_numbers
    masses = {'C': 12.011, 'H': 1.008}
    syms = sum([n * [s] for s, n in els], [])
    rng = np.random.RandomState(seed)
    mass = sum(masses[s] for s in syms)
    a0 = (1e24 * mass / (density * 6.02214076e23)) ** (1. / 3)
    assert all(s in atomic_numbers for s in syms)
    return S.Atoms(syms, rng.rand(len(syms), 3) * a0, [a0, a0, a0], True)


def _aC_small():
    d = load_npz('aC_small.npz')
    return S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)


def table():
    sq2 = np.sqrt(2.0)
    dia = lambda el, **kw: S.diamond(el, kw.pop('a0'), (sx, sx, sx))
    rows = [
        (ab.Harmonic, dict(el1='He', el2='He', k=1.0, r0=1.0, cutoff=1.5),
         [('fcc-He', S.fcc('He', sq2, (sx, sx, sx)))]),
        (ab.r6, dict(el1='Si', el2='Si', A=1.0, r0=1.0, cutoff=5.0),
         [('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx)))]),
        (ab.LJCut, dict(el1='He', el2='He', epsilon=10.2, sigma=2.28, cutoff=5.0, shift=True),
         [dict(name='fcc-He', struct=S.fcc('He', 3.5, (sx, sx, sx)), mask=True, rattle=0.1)]),
        (ab.Brenner, ab.Erhart_PRB_71_035211_SiC,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx))), ('dia-Si-C', S.b3(['Si', 'C'], 4.3596, (sx, sx, sx)))]),
        (ab.BrennerScr, ab.Erhart_PRB_71_035211_SiC__Scr,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx))), ('dia-Si-C', S.b3(['Si', 'C'], 4.3596, (sx, sx, sx)))]),
        (ab.Brenner, ab.Henriksson_PRB_79_114107_FeC,
         [dict(name='dia-C', struct=S.diamond('C', 3.57, (sx, sx, sx)), mask=True),
          dict(name='a-C', struct=_aC_small(), mask=True),
          dict(name='bcc-Fe', struct=S.bcc('Fe', 2.87, (sx, sx, sx)), mask=True),
          dict(name='fcc-Fe', struct=S.fcc('Fe', 3.6, (sx, sx, sx)), mask=True),
          dict(name='sc-Fe', struct=S.sc('Fe', 2.4, (sx, sx, sx)), mask=True),
          dict(name='B1-Fe-C', struct=S.b1(['Fe', 'C'], 3.9, (sx, sx, sx)), mask=True),
          dict(name='B3-Fe-C', struct=S.b3(['Fe', 'C'], 4.0, (sx, sx, sx)), mask=True)]),
        (ab.Kumagai, ab.Kumagai_CompMaterSci_39_457_Si, [('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx)))]),
        (ab.KumagaiScr, ab.Kumagai_CompMaterSci_39_457_Si__Scr, [('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx)))]),
        (ab.Tersoff, ab.Tersoff_PRB_39_5566_Si_C,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx))), ('dia-Si-C', S.b3(['Si', 'C'], 4.3596, (sx, sx, sx)))]),
        (ab.TersoffScr, ab.Tersoff_PRB_39_5566_Si_C__Scr,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('dia-Si', S.diamond('Si', 5.43, (sx, sx, sx))), ('dia-Si-C', S.b3(['Si', 'C'], 4.3596, (sx, sx, sx)))]),
        (ab.Rebo2, None,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('random-C-H', random_solid([('C', 50), ('H', 10)], 3.0))]),
        (ab.Rebo2Scr, None,
         [('dia-C', S.diamond('C', 3.57, (sx, sx, sx))), ('a-C', _aC_small()),
          ('random-C-H', random_solid([('C', 50), ('H', 10)], 3.0))]),
        (ab.TabulatedEAM, dict(funcfl=load_npz('au_u3_funcfl.npz')),
         [dict(name='fcc-Au', struct=S.fcc('Au', 4.08, (sx, sx, sx)), rattle=0.1)]),
        (ab.TabulatedAlloyEAM, dict(setfl=load_npz('au_grochola_setfl.npz')),
         [dict(name='fcc-Au', struct=S.fcc('Au', 4.08, (sx, sx, sx)), rattle=0.1, mask=True)]),
    ]
    del dia
    return rows


def _ids():
    return ['%s-%d' % (r[0].__name__, n) for n, r in enumerate(table())]


@pytest.mark.parametrize('row', range(len(_ids())), ids=_ids())
def test_forces_and_virial(row):
    pot, par, mats = table()[row]
    par = {k: v for k, v in (par or {}).items() if k != '__ref__'}
    c = pot(**par)
    rng = np.random.RandomState(row)
    for imat in mats:
        rattle, mask = 0.5, False
        if isinstance(imat, tuple):
            name, a = imat
        else:
            name, a = imat['name'], imat['struct']
            rattle, mask = imat.get('rattle', rattle), imat.get('mask', mask)
        a.positions = a.positions + 0.1
        a.calc = c
        masks = [None]
        if mask:
            masks += [(rng.randint(0, len(a), size=len(a)) < len(a) / 2).astype(np.int32),
                      (rng.randint(0, len(a), size=len(a)) < len(a) / 4).astype(np.int32)]
        for state in ('equilibrium', 'distorted'):
            for m in masks:
                c.set_mask(m)
                ffd, f0, maxdf = forces(a, dx=dx)
                assert abs(maxdf) < tol, (pot.__name__, name, state, 'forces', maxdf)
                sfd, s0, maxds = virial(a, de=dx)
                assert abs(maxds) < tol, (pot.__name__, name, state, 'virial', maxds)
            a.rattle(rattle, seed=row + 1)
        c.set_mask(None)
