"""Rebo2Scr on the GPU vs the oracle (1e-10 relative), through the C ABI.

The per-atom logic is the source that tests/test_emu_rebo2_scr.py also runs on the CPU against the
oracle; first green hardware run: round 2, first GPU call (124 passed with all fences lifted).
"""
import os

import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S

pytestmark = [pytest.mark.gpu]
RTOL = 1e-10


def _both(atoms, per_bond=False, **kw):
    p = native.from_atoms(atoms)
    nl = native.Neighbors(1000)
    pot = native.Rebo2Scr(**kw)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    rb = oracle.Rebo2Scr(**kw)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rb.cutoff(atoms.symbols), 1000)
    o = rb.energy_and_forces(atoms.positions, atoms.cell, onl, rb.ktyp(atoms.symbols), per_at=True,
                             per_bond=per_bond)
    return g, o


def _check(g, o, per_bond=False):
    e, f, w, epa, epb, fpb, wpa, wpb = g
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
    wscale = max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
    assert np.abs(w - o['wpot']).max() <= RTOL * wscale
    assert np.abs(epa - o['epot_per_at']).max() <= RTOL * max(1.0, np.abs(o['epot_per_at']).max())
    assert np.abs(wpa - o['wpot_per_at']).max() <= RTOL * max(1.0, np.abs(o['wpot_per_at']).max())
    if per_bond:
        n = len(epb)
        assert np.abs(epb - o['epot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['epot_per_bond']).max())
        assert np.abs(fpb - o['f_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['f_per_bond']).max())
        assert np.abs(wpb - o['wpot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['wpot_per_bond']).max())


def test_diamond():
    a = S.diamond('C', 3.566, (3, 3, 3))
    g, o = _both(a)
    _check(g, o)
    a.rattle(0.15, seed=1)
    g, o = _both(a)
    _check(g, o)


def test_amorphous_carbon(aC, aC_small):
    g, o = _both(aC_small, per_bond=True)
    _check(g, o, per_bond=True)
    g, o = _both(aC)
    _check(g, o)


def test_hydrocarbon_solid():
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (2, 2, 2))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    g, o = _both(a, per_bond=True)
    _check(g, o, per_bond=True)


def test_tiny_cell_self_images():
    a = S.diamond('C', 3.566, (1, 1, 1))
    a.rattle(0.05, seed=2)
    g, o = _both(a)
    _check(g, o)


def test_screening_table_grows():
    # low-density random carbon has many partially screened bonds per atom: the library doubles the
    # per-atom screening table and repeats the pass
    rng = np.random.RandomState(7)
    n, box = 200, 13.5
    pos = []
    while len(pos) < n:
        q = rng.uniform(0, box, 3)
        if all(np.linalg.norm((q - r + box / 2) % box - box / 2) > 1.25 for r in pos):
            pos.append(q)
    a = S.Atoms(['C'] * n, np.array(pos), [box, box, box], True)
    g, o = _both(a)
    _check(g, o)


def test_other_screening_parameters(aC_small):
    g, o = _both(aC_small, Cmin=0.8, Cmax=2.6, cc_nc_r2=3.4)
    _check(g, o)


def test_alt_dihedral(aC_small):
    """with_dihedral switches on the dihedral term of the screened build (ALT_DIHEDRAL,
    bop_kernel_rebo2.f90:2089-2371); the same per-atom source runs against the oracle on the CPU in
    tests/test_emu_rebo2_scr.py::test_alt_dihedral, the oracle's term is checked by finite differences in
    tests/test_oracle_kat.py"""
    g, o = _both(aC_small, per_bond=True, with_dihedral=True)
    _check(g, o, per_bond=True)
    g0, _ = _both(aC_small)
    assert abs(g[0] - g0[0]) > 0.1
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (3, 3, 3))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=2)
    g, o = _both(a, with_dihedral=True)
    _check(g, o)
    a = S.diamond('C', 3.566, (1, 1, 1))
    a.rattle(0.1, seed=5)
    g, o = _both(a, with_dihedral=True)
    _check(g, o)
