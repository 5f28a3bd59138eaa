"""The reference's tests/test_forces_and_virial.py driven through the CPU oracle.

Same table and procedure as tests/test_gpu_forces_and_virial.py (shift by 0.1, equilibrium and
rattled, random masks, FD helpers of atomistica_b200.tests, dx = 1e-6, tol = 1e-2); the calculator
behind the atoms object is an adapter around the oracle.  This pins the oracle on the reference's own
finite-difference procedure (including its 0.5 A rattle) and checks the table and the helpers that the
GPU version of this test uses.
"""
import numpy as np
import pytest

import oracle
import atomistica_b200 as ab
import test_gpu_forces_and_virial as G
import test_oracle_kat as K
from atomistica_b200 import parameters as P
from atomistica_b200.tests import test_forces as forces, test_virial as virial


class OracleCalc:
    """ASE calculator protocol around one of the oracle closures of test_oracle_kat.py"""

    def __init__(self, fn, masks=True):
        self.fn, self.mask, self.masks = fn, None, masks

    def set_mask(self, mask):
        self.mask = mask

    def _eval(self, atoms):
        kw = {} if self.mask is None else dict(mask=np.ascontiguousarray(self.mask, dtype=np.int32))
        return self.fn(atoms, **kw)

    def get_potential_energy(self, atoms):
        return self._eval(atoms)['epot']

    def get_forces(self, atoms):
        return self._eval(atoms)['f']

    def get_stress(self, atoms):
        w = self._eval(atoms)['wpot'] / atoms.get_volume()
        return np.array([w[0, 0], w[1, 1], w[2, 2], (w[1, 2] + w[2, 1]) / 2, (w[0, 2] + w[2, 0]) / 2,
                         (w[0, 1] + w[1, 0]) / 2])


def oracle_calculator(pot, par):
    par = {k: v for k, v in (par or {}).items() if k != '__ref__'}
    if pot in (ab.Tersoff, ab.Kumagai, ab.Brenner):
        return OracleCalc(K.bop_calc(pot.__name__, par or None))
    if pot in (ab.TersoffScr, ab.KumagaiScr, ab.BrennerScr):
        return OracleCalc(K.bop_scr_calc(pot.__name__[:-3], par or None))
    if pot is ab.Rebo2:
        return OracleCalc(K.rebo2_calc())
    if pot is ab.Rebo2Scr:
        return OracleCalc(K.rebo2_scr_calc())
    if pot is ab.TabulatedAlloyEAM:
        return OracleCalc(K.eam_calc(par['setfl']))
    if pot is ab.TabulatedEAM:
        return OracleCalc(K.funcfl_calc(par['funcfl']))
    kind = {ab.LJCut: oracle.PAIR_LJCUT, ab.Harmonic: oracle.PAIR_HARMONIC, ab.r6: oracle.PAIR_R6}[pot]
    names = {ab.LJCut: ('epsilon', 'sigma', 'cutoff'), ab.Harmonic: ('k', 'r0', 'cutoff'),
             ab.r6: ('A', 'r0', 'cutoff')}[pot]
    return OracleCalc(K.pair_calc(kind, [par[n] for n in names], par['cutoff'], shift=par.get('shift', False),
                                  el1=par.get('el1', '*'), el2=par.get('el2', '*')))


@pytest.mark.parametrize('row', range(len(G._ids())), ids=G._ids())
def test_forces_and_virial_oracle(row):
    pot, par, mats = G.table()[row]
    c = oracle_calculator(pot, par)
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
                ffd, f0, maxdf = forces(a, dx=G.dx)
                assert abs(maxdf) < G.tol, (pot.__name__, name, state, 'forces', maxdf)
                sfd, s0, maxds = virial(a, de=G.dx)
                assert abs(maxds) < G.tol, (pot.__name__, name, state, 'virial', maxds)
            a.rattle(rattle, seed=row + 1)
        c.set_mask(None)
