"""TabulatedAlloyEAM on the GPU vs the oracle (tolerance 1e-10 relative, BASELINE.json)."""
import os

import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _close(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300) if scale is None else scale
    return np.abs(a - b).max() <= RTOL * scale


def _both(atoms, setfl, mask=None, per_at=True):
    p = native.from_atoms(atoms)
    nl = native.Neighbors(200)
    pot = native.TabulatedAlloyEAM(setfl=setfl)
    pot.bind_to(p, nl)
    e, f, w, epa, _, _, wpa, _ = pot.energy_and_forces(p, nl, mask=mask, epot_per_at=per_at, wpot_per_at=per_at)
    eam = oracle.EAM(setfl)
    onl = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, eam.cutoff, 200)
    o = eam.energy_and_forces(atoms.positions, atoms.cell, onl, eam.eldb(atoms.symbols), mask=mask, per_at=per_at)
    return (e, f, w, epa, wpa), o


def _check(g, o, per_at=True):
    e, f, w, epa, wpa = g
    fscale = max(np.abs(o['f']).max(), 1.0)
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert _close(f, o['f'], fscale)
    # wpot = -dE/d(strain): cancelling sum of O(|E|) pair terms
    wscale = max(np.abs(o['wpot']).max(), 1.0, abs(o['epot']))
    assert _close(w, o['wpot'], wscale)
    if per_at:
        assert _close(epa, o['epot_per_at'])
        assert _close(wpa, o['wpot_per_at'], max(np.abs(o['wpot_per_at']).max(), 1.0))


def test_cu_perfect(cu_setfl):
    g, o = _both(S.fcc('Cu', 3.615, (5, 5, 5)), cu_setfl)
    _check(g, o)
    assert abs(g[0] / 500 + 3.54) < 0.01


def test_cu_rattled(cu_setfl):
    a = S.fcc('Cu', 3.615, (6, 5, 4))
    a.rattle(0.1, seed=3)
    g, o = _both(a, cu_setfl)
    _check(g, o)


def test_au_rattled(au_setfl):
    a = S.fcc('Au', 4.07, (4, 4, 4))
    a.rattle(0.1, seed=4)
    g, o = _both(a, au_setfl)
    _check(g, o)


def test_mask_additivity(au_setfl):
    # tests/test_mask.py:35-81: mask + complement == unmasked
    a = S.fcc('Au', 4.07, (4, 4, 4))
    a.rattle(0.1, seed=5)
    rng = np.random.RandomState(6)
    mask = (rng.rand(len(a)) > 0.5).astype(np.int32)
    g0, o0 = _both(a, au_setfl)
    g1, o1 = _both(a, au_setfl, mask=mask)
    g2, o2 = _both(a, au_setfl, mask=1 - mask)
    _check(g1, o1)
    _check(g2, o2)
    assert abs(g1[0] + g2[0] - g0[0]) < 1e-6
    assert np.abs(g1[1] + g2[1] - g0[1]).max() < 1e-6
    assert np.abs(g1[2] + g2[2] - g0[2]).max() < 1e-6


def test_unknown_element_is_ignored(cu_setfl):
    a = S.fcc('Cu', 3.615, (4, 4, 4))
    a.rattle(0.05, seed=7)
    a.symbols[5] = 'Si'
    a.symbols[77] = 'Si'
    g, o = _both(a, cu_setfl)
    _check(g, o)
    assert np.all(g[1][5] == 0.0)


def test_compressed_extrapolation(cu_setfl):
    # tests/test_eam_special_cases.py:51-77: embedding function extrapolates beyond rho_max
    from conftest import load_npz
    from atomistica_b200.structures import Atoms
    for name in ('eam_crash1.npz', 'eam_crash2.npz'):
        d = load_npz(name)
        a = Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
        g, o = _both(a, cu_setfl)
        _check(g, o)
    a.set_cell(a.cell * 0.5, scale_atoms=True)
    g, o = _both(a, cu_setfl)
    _check(g, o)


def test_calculator_interface(cu_setfl):
    from atomistica_b200 import TabulatedAlloyEAM
    a = S.fcc('Cu', 3.615, (4, 4, 4))
    a.calc = TabulatedAlloyEAM(setfl=cu_setfl)
    e1 = a.get_potential_energy()
    f = a.get_forces()
    assert f.shape == (len(a), 3)
    assert np.abs(f).max() < 1e-10
    s = a.get_stress()
    assert s.shape == (6,)
    a.positions[0, 0] += 0.1
    e2 = a.get_potential_energy()
    assert e2 > e1


def test_calculator_verlet_shell_reuses_list(cu_setfl):
    """library-mode Verlet shell: forces stay within tolerance of the oracle while the list is reused,
    and the list is rebuilt once an atom has moved verlet_shell/2"""
    from atomistica_b200 import TabulatedAlloyEAM
    a = S.fcc('Cu', 3.615, (6, 6, 6))
    a.rattle(0.05, seed=11)
    calc = TabulatedAlloyEAM(setfl=cu_setfl, verlet_shell=0.5)
    eam = oracle.EAM(cu_setfl)
    rng = np.random.RandomState(3)
    for it in range(6):
        f = calc.get_forces(a)
        e = calc.results['energy']
        onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff, 200)
        o = eam.energy_and_forces(a.positions, a.cell, onl, eam.eldb(a.symbols))
        assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
        assert _close(f, o['f'], max(np.abs(o['f']).max(), 1.0))
        a.positions = a.positions + rng.uniform(-0.04, 0.04, size=a.positions.shape)
    builds, reused = calc.nl.counters()
    assert builds >= 1 and reused >= 2 and builds + reused == 6
    a.positions[7] += 0.3                       # beyond verlet_shell/2: must rebuild
    calc.get_forces(a)
    assert calc.nl.counters()[0] == builds + 1


def test_full_size_replication_invariance(cu_setfl):
    """BASELINE config C2 size (40^3 cells = 256 000 atoms): a rattled periodic 8^3 block replicated
    5x5x5 has the same environments as the block itself, so E = 125 E_block and the forces tile --
    which carries the oracle parity of the block (checked here too) to the full size.  Plus sum(f) = 0."""
    blk = S.fcc('Cu', 3.615, (8, 8, 8))
    blk.rattle(0.08, seed=21)
    g, o = _both(blk, cu_setfl, per_at=False)
    _check(g, o, per_at=False)
    big = blk.repeat(5)
    assert len(big) == 256000
    p = native.from_atoms(big)
    nl = native.Neighbors(200)
    pot = native.TabulatedAlloyEAM(setfl=cu_setfl)
    pot.bind_to(p, nl)
    e, f, w = pot.energy_and_forces(p, nl)[:3]
    assert abs(e - 125 * o['epot']) <= RTOL * abs(125 * o['epot'])
    fscale = max(np.abs(o['f']).max(), 1.0)
    assert np.abs(f.reshape(125, -1, 3) - o['f'][None]).max() <= RTOL * fscale
    assert np.abs(f.sum(axis=0)).max() <= 1e-9 * fscale * np.sqrt(len(big))
    wscale = max(np.abs(o['wpot']).max(), 1.0, abs(o['epot']))
    assert np.abs(w - 125 * o['wpot']).max() <= RTOL * 125 * wscale


def test_store_outputs_mode(cu_setfl):
    """C-ABI output modes: default ADDS into f / epot_per_at (reference semantics), store mode
    writes them; the calculator's alternating force buffers keep the previous result valid"""
    a = S.fcc('Cu', 3.615, (5, 5, 5))
    a.rattle(0.08, seed=31)
    g, o = _both(a, cu_setfl)
    p = native.from_atoms(a)
    nl = native.Neighbors(200)
    pot = native.TabulatedAlloyEAM(setfl=cu_setfl)
    pot.bind_to(p, nl)
    f = np.full((len(a), 3), 7.0)
    pot.energy_and_forces(p, nl, forces=f)
    assert _close(f - 7.0, o['f'], max(np.abs(o['f']).max(), 1.0) * 10)      # added on top of the 7.0
    pot.set_store_outputs(True)
    f = np.full((len(a), 3), 7.0)
    e, _, w, epa = pot.energy_and_forces(p, nl, forces=f, epot_per_at=True)[:4]
    assert _close(f, o['f'], max(np.abs(o['f']).max(), 1.0))                  # stored
    assert abs(e - o['epot']) <= RTOL * abs(o['epot']) and _close(epa, o['epot_per_at'])

    from atomistica_b200 import TabulatedAlloyEAM
    calc = TabulatedAlloyEAM(setfl=cu_setfl)
    f1 = calc.get_forces(a)
    keep = f1.copy()
    b = a.copy()
    b.rattle(0.05, seed=32)
    f2 = calc.get_forces(b)
    assert np.array_equal(f1, keep)                       # not clobbered by the next call
    assert _close(f1, o['f'], max(np.abs(o['f']).max(), 1.0))
    g2, o2 = _both(b, cu_setfl)
    assert _close(f2, o2['f'], max(np.abs(o2['f']).max(), 1.0))


def test_funcfl_tabulated_eam(au_funcfl):
    """TabulatedEAM (funcfl, pair term Z(r)^2/r) vs the oracle; oracle pinned by the reference's
    fcc-Au bulk row (tests/test_oracle_kat.py)"""
    eam = oracle.EAMFuncfl(au_funcfl)
    for rattle, size in ((0.0, (3, 3, 3)), (0.1, (4, 4, 4)), (0.2, (2, 2, 2))):
        a = S.fcc('Au', 4.08, size)
        if rattle:
            a.rattle(rattle, seed=41)
        p = native.from_atoms(a)
        nl = native.Neighbors(200)
        pot = native.TabulatedEAM(funcfl=au_funcfl)
        pot.bind_to(p, nl)
        e, f, w, epa = pot.energy_and_forces(p, nl, epot_per_at=True)[:4]
        onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff, 200)
        o = eam.energy_and_forces(a.positions, a.cell, onl, a.symbols, per_at=True)
        assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
        assert _close(f, o['f'], max(np.abs(o['f']).max(), 1.0))
        assert _close(w, o['wpot'], max(np.abs(o['wpot']).max(), 1.0, abs(o['epot'])))
        assert _close(epa, o['epot_per_at'])
    assert abs(e / len(a)) > 3.0
    with pytest.raises(RuntimeError):
        pot.energy_and_forces(p, nl, wpot_per_at=True)


@pytest.mark.parametrize('lanes,unroll', [(4, 2), (8, 2), (16, 1)])
def test_consecutive_lane_mapping(cu_setfl, monkeypatch, lanes, unroll):
    # the variant the wavefront model of benchmarks/model_gather_wavefronts.py suggests: same sums in
    # another order.  per_at=False keeps the call on the fast kernels (the per-atom virial uses the
    # generic ones)
    monkeypatch.setenv('ATX_EAM_MAP', '1')
    monkeypatch.setenv('ATX_EAM_LANES', str(lanes))
    monkeypatch.setenv('ATX_EAM_UNROLL', str(unroll))
    a = S.fcc('Cu', 3.615, (6, 5, 4))
    a.rattle(0.1, seed=3)
    p = native.from_atoms(a)
    nl = native.Neighbors(200)
    pot = native.TabulatedAlloyEAM(setfl=cu_setfl)
    pot.bind_to(p, nl)
    e, f, w, epa, _, _, _, _ = pot.energy_and_forces(p, nl, epot_per_at=True)
    eam = oracle.EAM(cu_setfl)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff, 200)
    o = eam.energy_and_forces(a.positions, a.cell, onl, eam.eldb(a.symbols), per_at=True)
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert _close(f, o['f'], max(np.abs(o['f']).max(), 1.0))
    assert _close(w, o['wpot'], max(np.abs(o['wpot']).max(), 1.0, abs(o['epot'])))
    assert _close(epa, o['epot_per_at'])
