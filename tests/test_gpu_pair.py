"""Pair potentials (LJCut, Harmonic, DoubleHarmonic) on the GPU vs the oracle (1e-10 relative)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10

CASES = dict(
    LJCut=(native.LJCut, oracle.PAIR_LJCUT, ('epsilon', 'sigma', 'cutoff')),
    Harmonic=(native.Harmonic, oracle.PAIR_HARMONIC, ('k', 'r0', 'cutoff')),
    DoubleHarmonic=(native.DoubleHarmonic, oracle.PAIR_DOUBLE_HARMONIC, ('k1', 'r1', 'k2', 'r2', 'cutoff')),
    BornMayer=(native.BornMayer, oracle.PAIR_BORN_MAYER, ('A', 'rho', 'cutoff')),
    r6=(native.r6, oracle.PAIR_R6, ('A', 'r0', 'cutoff')),
)


def _both(name, atoms, par, shift=False, el1='*', el2='*', mask=None):
    cls, okind, keys = CASES[name]
    p = native.from_atoms(atoms)
    nl = native.Neighbors(400)
    pot = cls(element1=el1, element2=el2, **par) if name == 'BornMayer' else cls(el1=el1, el2=el2, shift=shift, **par)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=True, wpot_per_at=True)
    if name == 'BornMayer':
        assert not g[3].any() and not g[6].any() and not g[2].any()      # reference: energy and forces only
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, par['cutoff'], 400)
    o = oracle.pair_energy_and_forces(okind, [par[k] for k in keys], atoms.positions, atoms.cell, onl, atoms.symbols,
                                      el1=el1, el2=el2, shift=shift, mask=mask, per_at=True)
    return g, o


def _check_ef(g, o):
    assert abs(g[0] - o['epot']) <= RTOL * max(abs(o['epot']), 1.0)
    assert np.abs(g[1] - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())


def _check(g, o):
    e, f, w, epa, _, _, wpa, _ = g
    escale = max(abs(o['epot']), np.abs(o['epot_per_at']).sum(), 1.0)
    assert abs(e - o['epot']) <= RTOL * escale
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
    assert np.abs(w - o['wpot']).max() <= RTOL * max(1.0, np.abs(o['wpot']).max(), escale)
    assert np.abs(epa - o['epot_per_at']).max() <= RTOL * max(1.0, np.abs(o['epot_per_at']).max())
    assert np.abs(wpa - o['wpot_per_at']).max() <= RTOL * max(1.0, np.abs(o['wpot_per_at']).max())


LJ = dict(epsilon=0.0104, sigma=3.40, cutoff=8.0)


def test_lj_argon():
    a = S.fcc('Ar', 5.3, (4, 4, 4))
    a.rattle(0.2, seed=11)
    for shift in (False, True):
        g, o = _both('LJCut', a, LJ, shift=shift)
        _check(g, o)


def test_lj_mask_and_filters():
    a = S.fcc('Ar', 5.3, (3, 3, 3))
    a.rattle(0.2, seed=14)
    for i in range(0, len(a), 3):
        a.symbols[i] = 'Kr'
    rng = np.random.RandomState(4)
    mask = (rng.rand(len(a)) > 0.5).astype(np.int32)
    for el1, el2 in (('*', '*'), ('Ar', 'Kr'), ('Kr', 'Kr'), ('Ar,Kr', 'Ar')):
        g, o = _both('LJCut', a, LJ, el1=el1, el2=el2, mask=mask)
        _check(g, o)


def test_lj_self_images():
    # cutoff longer than the cell: i == j entries with a shift (weight w_i in the reference)
    a = S.fcc('Au', 4.08, (2, 2, 2))
    a.rattle(0.05, seed=2)
    rng = np.random.RandomState(1)
    mask = (rng.rand(len(a)) > 0.5).astype(np.int32)
    par = dict(epsilon=1.0, sigma=1.0, cutoff=6.0)            # tests/test_mask.py:76-79
    g, o = _both('LJCut', a, par, el1='Au', el2='Au')
    _check(g, o)
    g, o = _both('LJCut', a, par, el1='Au', el2='Au', mask=mask)
    _check(g, o)


def test_harmonic_and_double_harmonic():
    a = S.fcc('He', np.sqrt(2.0), (4, 4, 4))
    a.rattle(0.05, seed=12)
    g, o = _both('Harmonic', a, dict(k=1.0, r0=1.0, cutoff=1.3), shift=True)
    _check(g, o)
    a = S.sc('He', 1.0, (5, 5, 5))
    a.rattle(0.03, seed=13)
    par = dict(k1=1.0, r1=1.0, k2=1.0, r2=np.sqrt(2.0), cutoff=1.6)
    g, o = _both('DoubleHarmonic', a, par)
    _check(g, o)
    # 1x1x1 and 2x2x2 cells: every neighbour is a periodic image (self images for 1x1x1)
    for n in (1, 2):
        b = S.sc('He', 1.0, (n, n, n))
        b.positions += 0.01 * np.arange(3 * len(b)).reshape(-1, 3) / len(b)
        _check(*_both('DoubleHarmonic', b, par))
        _check(*_both('Harmonic', b, dict(k=1.0, r0=1.0, cutoff=1.3)))


def test_no_mask_for_harmonic():
    a = S.sc('He', 1.0, (3, 3, 3))
    p = native.from_atoms(a)
    nl = native.Neighbors(100)
    pot = native.Harmonic()
    pot.bind_to(p, nl)
    with pytest.raises(RuntimeError):
        pot.energy_and_forces(p, nl, mask=np.ones(len(a), dtype=np.int32))


def test_calculator_interface():
    from atomistica_b200 import LJCut
    a = S.fcc('Ar', 5.3, (3, 3, 3))
    a.rattle(0.1, seed=5)
    a.calc = LJCut(**LJ)
    e, f = a.get_potential_energy(), a.get_forces()
    g, o = _both('LJCut', a, LJ)
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())


def test_r6():
    a = S.fcc('Ar', 5.3, (3, 3, 3))
    a.rattle(0.2, seed=15)
    for i in range(0, len(a), 4):
        a.symbols[i] = 'Kr'
    for el1, el2 in (('*', '*'), ('Ar', 'Kr')):
        g, o = _both('r6', a, dict(A=50.0, r0=0.5, cutoff=7.0), el1=el1, el2=el2)
        _check(g, o)


def test_born_mayer_asymmetric_filters_and_self_images():
    a = S.b1(['Na', 'Cl'], 5.64, (2, 2, 2))
    a.rattle(0.1, seed=16)
    par = dict(A=1000.0, rho=0.3, cutoff=6.0)           # cutoff > half the cell: image entries of an atom with itself
    for el1, el2 in (('Na', 'Cl'), ('Cl', 'Na'), ('Na', 'Na'), ('*', 'Cl'), ('Na', '*'), ('*', '*')):
        g, o = _both('BornMayer', a, par, el1=el1, el2=el2)
        _check_ef(g, o)
