"""Tersoff / Kumagai / Brenner on the GPU vs the oracle (1e-10 relative, BASELINE.json)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, parameters as P, structures as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10

KINDS = dict(Tersoff=(native.Tersoff, oracle.TERSOFF), Kumagai=(native.Kumagai, oracle.KUMAGAI),
             Brenner=(native.Brenner, oracle.BRENNER))


def _both(kind, db, atoms, mask=None, per_bond=False, avgn=100):
    cls, okind = KINDS[kind]
    db = P.complete(kind, db)
    p = native.from_atoms(atoms)
    nl = native.Neighbors(avgn)
    pot = cls(db)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    present = [s for s in db['el'] if s in atoms.symbols]
    idx = [db['el'].index(s) for s in present]
    cutoff = max(db['r2'][P.pair_index(i, j, len(db['el']))] for i in idx for j in idx)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, cutoff, avgn)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in atoms.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(okind, db), atoms.positions, atoms.cell, onl, el, mask=mask,
                                     per_at=True, per_bond=per_bond)
    return g, o, onl


def _check(g, o, onl=None, per_bond=False):
    e, f, w, epa, epb, fpb, wpa, wpb = g
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
    # wpot = -dE/d(strain) is a sum of O(|E|) bond terms that cancel (almost completely in a relaxed
    # crystal): the relative tolerance refers to the energy scale of what is summed
    wscale = max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
    assert np.abs(w - o['wpot']).max() <= RTOL * wscale
    assert np.abs(epa - o['epot_per_at']).max() <= RTOL * max(1.0, np.abs(o['epot_per_at']).max())
    assert np.abs(wpa - o['wpot_per_at']).max() <= RTOL * max(1.0, np.abs(o['wpot_per_at']).max())
    if per_bond:
        n = len(epb)
        assert np.abs(epb - o['epot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['epot_per_bond']).max())
        assert np.abs(fpb - o['f_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['f_per_bond']).max())
        assert np.abs(wpb - o['wpot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['wpot_per_bond']).max())


@pytest.mark.parametrize('rattle', [0.0, 0.1])
def test_tersoff_si_c1(rattle):
    # BASELINE config C1: Si diamond 8x8x8, a0 5.432
    a = S.diamond('Si', 5.432, (8, 8, 8))
    a.positions += 0.1
    if rattle:
        a.rattle(rattle, seed=1)
    g, o, _ = _both('Tersoff', None, a)
    _check(g, o)
    if not rattle:
        assert abs(g[0] / len(a) + 4.6295950127) < 1e-9


def test_tersoff_sic_b3():
    a = S.b3(['Si', 'C'], 4.3596, (3, 3, 3))
    a.rattle(0.08, seed=2)
    g, o, _ = _both('Tersoff', None, a)
    _check(g, o)


def test_tersoff_amorphous_carbon(aC):
    g, o, _ = _both('Tersoff', None, aC)
    _check(g, o)


def test_kumagai_si():
    a = S.diamond('Si', 5.429, (4, 4, 4))
    g, o, _ = _both('Kumagai', None, a)
    _check(g, o)
    assert abs(g[0] / len(a) + 4.6299992839) < 1e-9
    a.rattle(0.15, seed=3)
    g, o, _ = _both('Kumagai', None, a)
    _check(g, o)


@pytest.mark.parametrize('db', ['Erhart_PRB_71_035211_SiC', 'Albe_PRB_65_195124_PtC',
                                'Henriksson_PRB_79_114107_FeC', 'Kioseoglou_PSSb_245_1118_AlN',
                                'Brenner_PRB_42_9458_C_I', 'Brenner_PRB_42_9458_C_II'])
def test_brenner_parameter_sets(db):
    par = getattr(P, db)
    els = par['el']
    if len(els) == 2:
        a = S.b3(els, 4.4, (3, 3, 3))
    else:
        a = S.diamond(els[0], 3.6, (3, 3, 3))
    a.rattle(0.1, seed=4)
    g, o, _ = _both('Brenner', par, a)
    _check(g, o)


def test_tersoff_bcn_three_elements():
    par = P.Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N
    a = S.b3(['B', 'N'], 3.7, (3, 3, 3))
    for i in range(0, len(a), 7):
        a.symbols[i] = 'C'
    a.rattle(0.1, seed=5)
    g, o, _ = _both('Tersoff', par, a)
    _check(g, o)


def test_mask(aC_small):
    # tests/test_mask.py: mask + complement == unmasked; maskfac in {0,1,2} (bop_kernel.f90:1095-1102)
    a = aC_small
    rng = np.random.RandomState(7)
    mask = (rng.rand(len(a)) > 0.5).astype(np.int32)
    g0, o0, _ = _both('Tersoff', None, a)
    g1, o1, _ = _both('Tersoff', None, a, mask=mask)
    g2, o2, _ = _both('Tersoff', None, a, mask=1 - mask)
    _check(g1, o1)
    _check(g2, o2)
    assert abs(g1[0] + g2[0] - g0[0]) < 1e-6
    assert np.abs(g1[1] + g2[1] - g0[1]).max() < 1e-6
    assert np.abs(g1[2] + g2[2] - g0[2]).max() < 1e-6


def test_per_bond(aC_small):
    g, o, onl = _both('Tersoff', None, aC_small, per_bond=True)
    _check(g, o, onl, per_bond=True)
    g, o, onl = _both('Brenner', None, S.b3(['Si', 'C'], 4.36, (2, 2, 2)), per_bond=True)
    _check(g, o, onl, per_bond=True)


def test_unknown_elements_skipped():
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.05, seed=8)
    a.symbols[3] = 'Cu'
    a.symbols[40] = 'Cu'
    g, o, _ = _both('Tersoff', None, a)
    _check(g, o)
    assert np.all(g[1][3] == 0.0)


def test_tiny_periodic_cell_and_surface():
    # 1x1x1 cell: an atom is its own periodic neighbour-of-neighbour
    a = S.diamond('Si', 5.432, (1, 1, 1))
    a.rattle(0.05, seed=9)
    g, o, _ = _both('Tersoff', None, a)
    _check(g, o)
    # tests/test_pbc.py: Si(100) surface energy 2.309 J/m^2 and pbc=[T,T,F] == doubled cell
    a = S.diamond('Si', 5.432, (2, 2, 2))
    sx, sy, sz = np.diag(a.cell)
    from atomistica_b200 import Tersoff
    a.calc = Tersoff()
    e1 = a.get_potential_energy()
    a.pbc[:] = [True, True, False]
    e2 = a.get_potential_energy()
    a.pbc[:] = True
    a.set_cell([sx, sy, 2 * sz])
    e3 = a.get_potential_energy()
    assert e2 == e3
    esurf = (e2 - e1) / (2 * sx * sy) * 16.021766208
    assert abs(esurf - 2.309) < 0.001


def test_large_lattice_property():
    # full-size property check (no oracle): 64^3 cells = 2.1M atoms, perfect lattice energy/atom and
    # vanishing forces
    a = S.diamond('Si', 5.432, (64, 64, 64))
    p = native.from_atoms(a)
    nl = native.Neighbors(20)
    pot = native.Tersoff()
    pot.bind_to(p, nl)
    e, f, w = pot.energy_and_forces(p, nl)[:3]
    assert abs(e / len(a) + 4.6295950127) < 1e-9
    assert np.abs(f).max() < 1e-9


@pytest.mark.parametrize('kind,a0', [('Tersoff', 5.432), ('Kumagai', 5.429)])
def test_queued_pass_for_overcoordinated_atoms(kind, a0):
    """10x10x10 Si with one interstitial: the bond table is sized for 4 bonds (>= 99.9 % of the atoms),
    the handful of over-coordinated atoms go through the queued deep-table pass; per-bond outputs
    included so that every slot written by either pass is compared"""
    a = S.diamond('Si', a0, (10, 10, 10))
    a.rattle(0.03, seed=7)
    pos = np.vstack([a.positions, [[0.5 * a0 + 3 * a0, 0.5 * a0 + 3 * a0, 0.5 * a0 + 3 * a0]]])
    b = S.Atoms(list(a.symbols) + ['Si'], pos, a.cell, True)
    g, o, onl = _both(kind, None, b, per_bond=True)
    _check(g, o, onl, per_bond=True)
