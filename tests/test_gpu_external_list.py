"""LAMMPS-style operation: the host supplies the neighbour list, ghosts are explicit atoms.

A periodic system is unfolded the way LAMMPS would hand it over -- owned atoms plus ghost images
within a 2 x cutoff shell, full lists for owned AND ghost atoms (REQ_FULL|REQ_GHOST), no periodic
shifts -- and the result must equal the ordinary periodic calculation: energy, virial, forces on
the owned atoms; ghost rows of the force array stay zero."""
import os

import numpy as np
import pytest

from atomistica_b200 import native, structures as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def unfold(atoms, cutoff, shell=2):
    """owned atoms + ghost images within shell*cutoff of the (orthorhombic) box, brute-force full lists"""
    L = np.diag(atoms.cell)
    pos, sym = [atoms.positions], [list(atoms.symbols)]
    w = shell * cutoff
    rng = [range(-int(np.ceil(w / L[k])), int(np.ceil(w / L[k])) + 1) for k in range(3)]
    for i in rng[0]:
        for j in rng[1]:
            for k in rng[2]:
                if (i, j, k) == (0, 0, 0):
                    continue
                p = atoms.positions + np.array([i, j, k]) * L
                m = np.all((p > -w) & (p < L + w), axis=1)
                pos.append(p[m])
                sym.append([s for s, t in zip(atoms.symbols, m) if t])
    pos = np.concatenate(pos)
    sym = sum(sym, [])
    nall = len(pos)
    lists = []
    for i in range(nall):
        d2 = ((pos - pos[i]) ** 2).sum(axis=1)
        nb = np.nonzero(d2 < cutoff * cutoff)[0]
        lists.append(nb[nb != i].astype(np.int32))
    big = S.Atoms(sym, pos, L + 10 * w, False)
    return big, lists


SKIN = 0.3      # the host's list carries a skin; the kernels apply the potentials' own cutoffs


def _compare(atoms, make_pot, cutoff, avgn=200, shell=2):
    cutoff = cutoff + SKIN
    p0 = native.from_atoms(atoms)
    nl0 = native.Neighbors(avgn)
    pot0 = make_pot()
    pot0.bind_to(p0, nl0)
    e0, f0, w0 = pot0.energy_and_forces(p0, nl0)[:3]

    big, lists = unfold(atoms, cutoff, shell)
    nloc = len(atoms)
    p = native.from_atoms(big)
    nl = native.Neighbors(avgn)
    pot = make_pot()
    pot.bind_to(p, nl)
    nl.set_external(p, nloc, np.arange(len(big)), lists)
    e, f, w = pot.energy_and_forces(p, nl)[:3]
    assert abs(e - e0) <= RTOL * abs(e0)
    fs = max(1.0, np.abs(f0).max())
    assert np.abs(f[:nloc] - f0).max() <= RTOL * fs
    assert np.all(f[nloc:] == 0.0)
    assert np.abs(w - w0).max() <= RTOL * max(1.0, np.abs(w0).max(), abs(e0))
    # the host moves atoms between reneighbourings: positions follow, the list is kept
    d = np.random.RandomState(1).normal(scale=0.01, size=(nloc, 3))
    atoms2 = atoms.copy()
    atoms2.positions = atoms.positions + d
    big2, _ = unfold(atoms2, cutoff, shell)
    if len(big2) == len(big) and np.abs(big2.positions - big.positions).max() < 0.1:
        p.coordinates[:, :] = big2.positions
        p.I_changed_positions()
        p0.coordinates[:, :] = atoms2.positions
        p0.I_changed_positions()
        e1 = pot.energy_and_forces(p, nl)[0]
        e2 = pot0.energy_and_forces(p0, nl0)[0]
        assert abs(e1 - e2) <= 1e-6 * abs(e2)      # pairs crossing the cutoff may differ: list not rebuilt


def test_eam_external(cu_setfl):
    a = S.fcc('Cu', 3.615, (4, 4, 4))
    a.rattle(0.05, seed=3)
    _compare(a, lambda: native.TabulatedAlloyEAM(setfl=cu_setfl), float(cu_setfl['cutoff']))


def test_eam_external_mask_and_per_atom_virial(cu_setfl):
    """masks and per-atom energies / virials of TabulatedAlloyEAM on a caller-supplied list (the generic
    kernels honour the owned / ghost roles): owned rows equal the periodic calculation"""
    a = S.fcc('Cu', 3.615, (4, 4, 4))
    a.rattle(0.05, seed=3)
    cutoff = float(cu_setfl['cutoff']) + SKIN
    mask = (np.random.RandomState(7).rand(len(a)) > 0.4).astype(np.int32)
    p0 = native.from_atoms(a)
    nl0 = native.Neighbors(200)
    pot0 = native.TabulatedAlloyEAM(setfl=cu_setfl)
    pot0.bind_to(p0, nl0)
    big, lists = unfold(a, cutoff, 2)
    nloc = len(a)
    # the mask of a ghost is the mask of the atom it is an image of
    L = np.diag(a.cell)
    frac = np.mod(big.positions / L, 1.0)
    frac0 = np.mod(a.positions / L, 1.0)
    owner = np.array([np.argmin(((np.abs(frac0 - fr) + 0.5) % 1.0 - 0.5).__abs__().sum(axis=1)) for fr in frac])
    assert np.array_equal(owner[:nloc], np.arange(nloc))
    p = native.from_atoms(big)
    nl = native.Neighbors(200)
    pot = native.TabulatedAlloyEAM(setfl=cu_setfl)
    pot.bind_to(p, nl)
    nl.set_external(p, nloc, np.arange(len(big)), lists)
    for m0 in (None, mask):
        m = None if m0 is None else m0[owner].astype(np.int32)
        e0, f0, w0, epa0, _, _, wpa0, _ = pot0.energy_and_forces(p0, nl0, mask=m0, epot_per_at=True, wpot_per_at=True)
        e, f, w, epa, _, _, wpa, _ = pot.energy_and_forces(p, nl, mask=m, epot_per_at=True, wpot_per_at=True)
        assert abs(e - e0) <= RTOL * abs(e0)
        assert np.abs(f[:nloc] - f0).max() <= RTOL * max(1.0, np.abs(f0).max())
        assert np.all(f[nloc:] == 0.0)
        assert np.abs(w - w0).max() <= RTOL * max(1.0, np.abs(w0).max(), abs(e0))
        assert np.abs(epa[:nloc] - epa0).max() <= RTOL * max(1.0, np.abs(epa0).max())
        assert np.abs(wpa[:nloc] - wpa0).max() <= RTOL * max(1.0, np.abs(wpa0).max())
        assert np.all(epa[nloc:] == 0.0)


def test_tersoff_external():
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.08, seed=4)
    _compare(a, native.Tersoff, 3.0, avgn=50)


def test_brenner_sic_external():
    a = S.b3(['Si', 'C'], 4.3596, (3, 3, 3))
    a.rattle(0.08, seed=5)
    _compare(a, native.Brenner, 2.96, avgn=50)


def test_lj_external():
    a = S.fcc('Ar', 5.3, (3, 3, 3))
    a.rattle(0.1, seed=6)
    _compare(a, lambda: native.LJCut(epsilon=0.0104, sigma=3.40, cutoff=6.0), 6.0)


def test_rebo2_external_refuses_per_bond_outputs():
    a = S.diamond('C', 3.566, (2, 2, 2))
    big, lists = unfold(a, 2.0)
    p = native.from_atoms(big)
    nl = native.Neighbors(50)
    pot = native.Rebo2()
    pot.bind_to(p, nl)
    nl.set_external(p, len(a), np.arange(len(big)), lists)
    with pytest.raises(RuntimeError):
        pot.energy_and_forces(p, nl, epot_per_bond=True)


def test_rebo2_external():
    # REBO2 reaches five bonds: ghosts within 5 bond cutoffs
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.6, (3, 3, 3))
    for i in rng.choice(len(a), len(a) // 4, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=4)
    _compare(a, native.Rebo2, 2.0, avgn=50, shell=5)
