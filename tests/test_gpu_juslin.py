"""Juslin (W-C-H) / Kuopanportti (Fe-C-H) on the GPU vs the oracle (1e-10 relative); the oracle is
pinned by the reference's bulk-property table (tests/test_oracle_kat.py)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, parameters as P, structures as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _both(db, atoms, mask=None, per_bond=False):
    db = P.complete_juslin(db)
    p = native.from_atoms(atoms)
    nl = native.Neighbors(200)
    pot = native.Juslin(db)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    idx = [db['el'].index(s) for s in set(atoms.symbols) if s in db['el']]
    cutoff = max(db['r2'][j + i * len(db['el'])] for i in idx for j in idx)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, cutoff, 200)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in atoms.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.JUSLIN, db), atoms.positions, atoms.cell, onl, el,
                                     mask=mask, per_at=True, per_bond=per_bond)
    return g, o


def _check(g, o, per_bond=False):
    e, f, w, epa, epb, fpb, wpa, wpb = g
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
    assert np.abs(w - o['wpot']).max() <= RTOL * max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
    assert np.abs(epa - o['epot_per_at']).max() <= RTOL * max(1.0, np.abs(o['epot_per_at']).max())
    assert np.abs(wpa - o['wpot_per_at']).max() <= RTOL * max(1.0, np.abs(o['wpot_per_at']).max())
    if per_bond:
        n = len(epb)
        assert np.abs(epb - o['epot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['epot_per_bond']).max())
        assert np.abs(fpb - o['f_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['f_per_bond']).max())
        assert np.abs(wpb - o['wpot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['wpot_per_bond']).max())


def test_bcc_tungsten():
    a = S.bcc('W', 3.165, (4, 4, 4))
    g, o = _both(None, a)
    _check(g, o)
    assert abs(g[0] / len(a) + 8.89) < 0.01          # cohesive energy of the paper
    a.rattle(0.1, seed=1)
    g, o = _both(None, a)
    _check(g, o)


def test_wch_mixture():
    a = S.b1(['W', 'C'], 4.38, (3, 3, 3))
    rng = np.random.RandomState(5)
    for i in rng.choice(len(a), 30, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.15, seed=9)
    g, o = _both(None, a, per_bond=True)
    _check(g, o, per_bond=True)


def test_hydrocarbon_solid_and_mask():
    a = S.diamond('C', 3.7, (3, 3, 3))
    rng = np.random.RandomState(6)
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    g, o = _both(None, a)
    _check(g, o)
    mask = (rng.rand(len(a)) > 0.5).astype(np.int32)
    g, o = _both(None, a, mask=mask)
    _check(g, o)


def test_fe_c_h():
    a = S.bcc('Fe', 2.87, (4, 4, 4))
    for i in (0, 7, 20, 33):
        a.symbols[i] = 'C'
    a.symbols[50] = 'H'
    a.rattle(0.1, seed=7)
    g, o = _both(P.Kuopanportti_CMS_111_525_FeCH, a)
    _check(g, o)


# ---- JuslinScr (juslin_scr.f90: the same module compiled with SCREENING) -----------------------------

def _wide_scr_db():
    """default JuslinScr database with outer / bond-order cutoffs beyond the inner one, so that bonds
    are really screened (the Fortran default has or = bor = r)"""
    db = P.complete_juslin_scr(None)
    for k in range(9):
        if db['r2'][k] > 0:
            db['or1'][k], db['or2'][k] = db['r2'][k] * 1.05, db['r2'][k] * 1.45
            db['bor1'][k], db['bor2'][k] = db['r2'][k] * 1.0, db['r2'][k] * 1.35
    return db


def _both_scr(db, atoms, mask=None, per_bond=False):
    db = P.complete_juslin_scr(db)
    p = native.from_atoms(atoms)
    nl = native.Neighbors(1000)
    pot = native.JuslinScr(db)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    # the list cutoff the device asked for: the largest of the pairs present (juslin_module.f90:379-395)
    nel = len(db['el'])
    idx = [db['el'].index(s) for s in set(atoms.symbols) if s in db['el']]
    m = max(max(db['r2'][k], db['or2'][k], db['bor2'][k]) for k in range(nel * nel))
    cutoff = max((db['Cmax'][j + i * nel] ** 2 / (4 * (db['Cmax'][j + i * nel] - 1))) ** 0.5 * m
                 for i in idx for j in idx)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, cutoff, 1000)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in atoms.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.JUSLIN, db), atoms.positions, atoms.cell, onl, el,
                                     scr=oracle.bop_scr_params(db), mask=mask, per_at=True, per_bond=per_bond)
    return g, o


def test_juslin_scr_default_database():
    a = S.b1(['W', 'C'], 4.38, (3, 3, 3))
    a.rattle(0.05, seed=2)
    g, o = _both_scr(None, a)
    _check(g, o)
    b = S.bcc('W', 3.165, (4, 4, 4))
    b.rattle(0.1, seed=1)
    g, o = _both_scr(None, b)
    _check(g, o)


def test_juslin_scr_screened_bonds_wch():
    db = _wide_scr_db()
    a = S.b1(['W', 'C'], 4.38, (3, 3, 3))
    rng = np.random.RandomState(5)
    for i in rng.choice(len(a), 30, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.15, seed=9)
    g, o = _both_scr(db, a, per_bond=True)
    _check(g, o, per_bond=True)
    g1, _ = _both_scr(None, a)
    assert abs(g[0] - g1[0]) > 1e-6              # the wider cutoffs matter
    b = S.bcc('W', 3.165, (3, 3, 3))
    b.rattle(0.15, seed=2)
    g, o = _both_scr(db, b)
    _check(g, o)


def test_juslin_scr_mask_and_hydrocarbon():
    db = _wide_scr_db()
    a = S.diamond('C', 3.7, (3, 3, 3))
    rng = np.random.RandomState(6)
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    g, o = _both_scr(db, a)
    _check(g, o)
    mask = (np.random.RandomState(4).rand(len(a)) > 0.5).astype(np.int32)
    g, o = _both_scr(db, a, mask=mask)
    _check(g, o)


def test_juslin_scr_calculator_and_md():
    import atomistica_b200 as ab
    a = S.bcc('W', 3.165, (4, 4, 4))
    a.rattle(0.05, seed=3)
    calc = ab.JuslinScr(db=_wide_scr_db())
    e = calc.get_potential_energy(a)
    g, o = _both_scr(_wide_scr_db(), a)
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
