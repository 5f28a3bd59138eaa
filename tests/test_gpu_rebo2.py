"""Rebo2 on the GPU vs the oracle (1e-10 relative)."""
import json
import os

import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _both(atoms, per_bond=False, **kw):
    p = native.from_atoms(atoms)
    nl = native.Neighbors(100)
    pot = native.Rebo2(**kw)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    okw = dict(kw)
    if 'dihedral' in okw:
        okw['with_dihedral'] = okw.pop('dihedral')
    rb = oracle.Rebo2(**okw)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rb.cutoff(atoms.symbols), 100)
    o = rb.energy_and_forces(atoms.positions, atoms.cell, onl, rb.ktyp(atoms.symbols), per_at=True,
                             per_bond=per_bond)
    return g, o


def _check(g, o, per_bond=False):
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


def test_diamond():
    a = S.diamond('C', 3.566, (3, 3, 3))
    g, o = _both(a)
    _check(g, o)
    assert abs(g[0] / len(a) + 7.370) < 0.005
    a.rattle(0.1, seed=1)
    g, o = _both(a)
    _check(g, o)


def test_amorphous_carbon(aC, aC_small):
    g, o = _both(aC_small, per_bond=True)
    _check(g, o, per_bond=True)
    g, o = _both(aC)          # triclinic, 4001 atoms (BASELINE config C3 before replication)
    _check(g, o)


def test_tiny_cell_self_images():
    a = S.diamond('C', 3.566, (1, 1, 1))
    a.rattle(0.05, seed=2)
    g, o = _both(a)
    _check(g, o)


def _molecule(name, vacuum=5.0):
    mols = json.load(open(os.path.join(GOLDEN, 'molecules.json')))
    m = mols[name]
    pos = np.array(m['positions'])
    pos -= pos.min(axis=0) - vacuum
    cell = pos.max(axis=0) + vacuum
    return S.Atoms(m['symbols'], pos, cell, True)


@pytest.mark.parametrize('name', ['cyclohexane', 'naphthalene', 'C2H', 'i-C4H9', 'propyne', '1,3-pentadiene'])
def test_hydrocarbons(name):
    a = _molecule(name)
    a.rattle(0.05, seed=3)
    g, o = _both(a, per_bond=True)
    _check(g, o, per_bond=True)


def test_random_CH_solid():
    # tests/test_forces_and_virial.py uses random C/H solids to exercise every table branch
    rng = np.random.RandomState(11)
    a = S.diamond('C', 3.7, (3, 3, 3))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.15, seed=4)
    g, o = _both(a)
    _check(g, o)


def test_dihedral(aC_small):
    g, o = _both(aC_small, dihedral=True)
    _check(g, o)
    g2, _ = _both(aC_small)
    assert abs(g[0] - g2[0]) > 1e-6      # the dihedral term is really on
    a = _molecule('1,3-pentadiene')
    a.rattle(0.05, seed=5)
    g, o = _both(a, dihedral=True)
    _check(g, o)


def test_no_mask_support(aC_small):
    p = native.from_atoms(aC_small)
    nl = native.Neighbors(100)
    pot = native.Rebo2()
    pot.bind_to(p, nl)
    with pytest.raises(RuntimeError):
        pot.energy_and_forces(p, nl, mask=np.ones(len(aC_small), dtype=np.int32))


def test_full_size_replication_invariance(aC):
    """BASELINE config C3: aC.cfg (4001 atoms, triclinic) replicated 5x5x5 = 500 125 atoms.  The
    replica has the environments of the periodic original, so E = 125 E_aC and the forces tile; the
    original is compared with the oracle in test_amorphous_carbon."""
    g, o = _both(aC)
    big = aC.repeat(5)
    assert len(big) == 500125
    p = native.from_atoms(big)
    nl = native.Neighbors(50)
    pot = native.Rebo2()
    pot.bind_to(p, nl)
    e, f = pot.energy_and_forces(p, nl)[:2]
    assert abs(e - 125 * o['epot']) <= RTOL * abs(125 * o['epot'])
    fscale = max(np.abs(o['f']).max(), 1.0)
    assert np.abs(f.reshape(125, -1, 3) - o['f'][None]).max() <= RTOL * fscale
    assert np.abs(f.sum(axis=0)).max() <= 1e-9 * fscale * np.sqrt(len(big))


@pytest.mark.parametrize('variant', ['1', '2', '3'])
def test_one_thread_per_bond_variant(aC_small, monkeypatch, variant):
    monkeypatch.setenv('ATX_REBO2_PERBOND', variant)      # 4 / 6 / 8 resident blocks per SM
    g, o = _both(aC_small, per_bond=True)
    _check(g, o, per_bond=True)
    g, o = _both(aC_small, dihedral=True)
    _check(g, o)
    a = S.diamond('C', 3.566, (1, 1, 1))
    a.rattle(0.05, seed=2)
    g, o = _both(a)
    _check(g, o)
