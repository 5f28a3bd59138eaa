"""TersoffScr / KumagaiScr / BrennerScr on the GPU vs the oracle (1e-10 relative).

The oracle's screened kernel is pinned by the reference's own tests (tests/test_oracle_kat.py:
bulk properties, finite differences, mask additivity, Si2 dimer)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, parameters as P, structures as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10

KINDS = dict(Tersoff=(native.TersoffScr, oracle.TERSOFF), Kumagai=(native.KumagaiScr, oracle.KUMAGAI),
             Brenner=(native.BrennerScr, oracle.BRENNER))


def _both(kind, db, atoms, mask=None, per_bond=False, avgn=1000):
    cls, okind = KINDS[kind]
    db = P.complete_scr(kind, db)
    p = native.from_atoms(atoms)
    nl = native.Neighbors(avgn)
    pot = cls(db)
    pot.bind_to(p, nl)
    g = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=True, wpot_per_at=True, epot_per_bond=per_bond,
                              f_per_bond=per_bond, wpot_per_bond=per_bond)
    cutoff = P.scr_cutoff(db)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, cutoff, avgn)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in atoms.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(okind, db), atoms.positions, atoms.cell, onl, el, mask=mask,
                                     per_at=True, per_bond=per_bond, scr=oracle.bop_scr_params(db))
    return g, o


def _check(g, o, per_bond=False):
    e, f, w, epa, epb, fpb, wpa, wpb = g
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
    wscale = max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
    assert np.abs(w - o['wpot']).max() <= RTOL * wscale
    assert np.abs(epa - o['epot_per_at']).max() <= RTOL * max(1.0, np.abs(o['epot_per_at']).max())
    assert np.abs(wpa - o['wpot_per_at']).max() <= RTOL * max(1.0, np.abs(o['wpot_per_at']).max(), abs(o['epot']) / len(f))
    if per_bond:
        n = len(epb)
        assert np.abs(epb - o['epot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['epot_per_bond']).max())
        assert np.abs(fpb - o['f_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['f_per_bond']).max())
        assert np.abs(wpb - o['wpot_per_bond'][:n]).max() <= RTOL * max(1.0, np.abs(o['wpot_per_bond']).max())


@pytest.mark.parametrize('kind,a0', [('Tersoff', 5.432), ('Kumagai', 5.429), ('Brenner', 5.429)])
def test_si_diamond(kind, a0):
    a = S.diamond('Si', a0, (3, 3, 3))
    g, o = _both(kind, None, a)
    _check(g, o)
    a.rattle(0.15, seed=1)
    g, o = _both(kind, None, a)
    _check(g, o)


@pytest.mark.parametrize('kind', ['Tersoff', 'Brenner'])
def test_sic_b3(kind):
    a = S.b3(['Si', 'C'], 4.3596, (3, 3, 3))
    a.rattle(0.1, seed=2)
    g, o = _both(kind, None, a)
    _check(g, o)


@pytest.mark.parametrize('kind', ['Tersoff', 'Brenner'])
def test_amorphous_carbon(kind, aC_small):
    # partially screened bonds with screening-neighbour derivatives; per-bond outputs
    g, o = _both(kind, None, aC_small, per_bond=True)
    _check(g, o, per_bond=True)


def test_amorphous_carbon_full(aC):
    g, o = _both('Tersoff', None, aC)
    _check(g, o)


def test_bcn_three_elements():
    a = S.b3(['B', 'N'], 3.7, (3, 3, 3))
    for i in (3, 40, 77):
        a.symbols[i] = 'C'
    a.rattle(0.08, seed=4)
    g, o = _both('Tersoff', P.Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N__Scr, a)
    _check(g, o)


def test_mask(aC_small):
    rng = np.random.RandomState(3)
    mask = (rng.rand(len(aC_small)) > 0.5).astype(np.int32)
    g, o = _both('Tersoff', None, aC_small, mask=mask)
    _check(g, o)


def test_compressed_and_stretched():
    # strongly compressed (many bonds, long screening lists) and stretched (bonds in the outer cutoff)
    for scale in (0.8, 1.15, 1.3):
        a = S.diamond('Si', 5.432 * scale, (3, 3, 3))
        a.rattle(0.1, seed=5)
        g, o = _both('Kumagai', None, a)
        _check(g, o)


def test_calculator_interface():
    from atomistica_b200 import TersoffScr
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.05, seed=6)
    a.calc = TersoffScr()
    e = a.get_potential_energy()
    f = a.get_forces()
    g, o = _both('Tersoff', None, a)
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())


def test_list_reuse_with_verlet_shell(aC_small):
    """calculator with a library-mode Verlet shell: the list (and the sizing of the screened tables)
    is reused while the atoms move; results stay on the oracle"""
    from atomistica_b200 import TersoffScr
    a = aC_small.copy()
    calc = TersoffScr(verlet_shell=0.8)
    rng = np.random.RandomState(2)
    db = P.complete_scr('Tersoff', None)
    par, scr = oracle.bop_params(oracle.TERSOFF, db), oracle.bop_scr_params(db)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    for it in range(4):
        f = calc.get_forces(a)
        e = calc.results['energy']
        onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, P.scr_cutoff(db), 1000)
        o = oracle.bop_energy_and_forces(par, a.positions, a.cell, onl, el, scr=scr)
        assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
        assert np.abs(f - o['f']).max() <= RTOL * max(1.0, np.abs(o['f']).max())
        a.positions = a.positions + rng.uniform(-0.1, 0.1, size=a.positions.shape)
    builds, reused = calc.nl.counters()
    assert reused >= 1
