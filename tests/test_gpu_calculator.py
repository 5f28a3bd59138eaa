"""Calculator front end (atomistica_b200.aseinterface): result arrays are private to the caller."""
import numpy as np
import pytest

import atomistica_b200 as ab
from atomistica_b200 import structures as S

pytestmark = pytest.mark.gpu


def test_forces_are_not_overwritten_while_held():
    """ase.calculators.calculator.Calculator.get_property returns copies; here the page-locked result
    buffers are recycled only when the caller dropped them (f_old aliased f_new)"""
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.05, seed=1)
    calc = ab.Tersoff()
    held, copies = [], []
    rng = np.random.RandomState(0)
    for k in range(12):                       # more than the pool holds: the overflow path copies
        a.positions += rng.normal(scale=0.01, size=a.positions.shape)
        f = calc.get_forces(a)
        calc.get_potential_energy(a)          # an intermediate call must not touch f either
        held.append(f)
        copies.append(f.copy())
    for f, c in zip(held, copies):
        assert np.array_equal(f, c)
    view = held[-1][:5]
    del held, f
    keep = view.copy()
    for k in range(20):                       # a VIEW keeps its buffer out of the pool as well
        a.positions += rng.normal(scale=0.01, size=a.positions.shape)
        calc.get_forces(a)
    assert np.array_equal(view, keep)
    assert len(calc._fbuf) <= calc.MAX_FORCE_BUFFERS


def test_stresses_are_voigt_wpot_per_at():
    """aseinterface.py:438-446: 'stresses' is the Voigt-ordered wpot_per_at, not divided by the volume;
    its sum is stress * volume"""
    a = S.diamond('Si', 5.432, (2, 2, 2))
    a.rattle(0.05, seed=2)
    calc = ab.Tersoff()
    s = calc.get_stresses(a)
    st = calc.get_stress(a)
    assert s.shape == (len(a), 6)
    assert np.abs(s.sum(axis=0) - st * a.get_volume()).max() < 1e-9 * max(1.0, np.abs(st * a.get_volume()).max())


def test_positions_buffer_is_used_in_place():
    """two calls with the same positions buffer make that buffer r_non_cyc itself (page-locked where it
    lies, no mirror copy); results equal those of a calculator that copies; switching to another array
    falls back to the mirror; an array that is already page-locked is accepted and left locked"""
    from atomistica_b200 import _lib as L
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.05, seed=3)
    calc = ab.Tersoff(verlet_shell=0.4)
    ref = ab.Tersoff(alias_positions=False)
    rng = np.random.RandomState(1)
    for k in range(5):
        a.positions += rng.normal(scale=0.01, size=a.positions.shape)      # in place: same buffer
        f = calc.get_forces(a)
        e = calc.get_potential_energy(a)
        b = a.copy()
        assert np.abs(f - ref.get_forces(b)).max() < 1e-10
        assert abs(e - ref.get_potential_energy(b)) < 1e-10 * abs(e)
        assert ref.particles._alias is None
    assert calc.particles._alias is a.positions and calc.particles._alias_registered
    assert calc.particles.coordinates is a.positions
    # another array: back to the mirror, and this one is adopted after it was seen twice
    a.positions = a.positions + 0.01
    f = calc.get_forces(a)
    assert calc.particles._alias is None
    assert np.abs(f - ref.get_forces(a.copy())).max() < 1e-10
    calc.get_forces(a)
    assert calc.particles._alias is a.positions
    # a non-contiguous view is never adopted
    big = np.zeros((len(a), 4))
    big[:, :3] = a.positions
    a.positions = big[:, :3]
    for k in range(3):
        f = calc.get_forces(a)
    assert calc.particles._alias is None
    assert np.abs(f - ref.get_forces(a.copy())).max() < 1e-10
    # memory that is page-locked already
    pinned = L.PinnedArray((len(a), 3))
    pinned.array[...] = big[:, :3]
    a.positions = pinned.array
    for k in range(3):
        a.positions += 0.001
        f = calc.get_forces(a)
    assert calc.particles._alias is pinned.array and not calc.particles._alias_registered
    assert np.abs(f - ref.get_forces(a.copy())).max() < 1e-10
    del calc                                    # unregisters nothing it did not register
    pinned.array[...] = 0.0


def test_rebo2_dimer_results_are_reproducible():
    """the reference's tests/test_io.py: energy, forces and stress of a C2 dimer in vacuum with Rebo2 and
    Rebo2Scr survive a round trip of the structure (ASE's trajectory file there; a copy of the Atoms here) --
    i.e. a fresh calculation on the same geometry gives the same numbers to 1e-10"""
    vac, dist_min = 8.0, 1.2
    pos = np.array([[0.0, 0.0, 0.0], [dist_min, 0.0, 0.0]]) + vac
    a = S.Atoms(['C', 'C'], pos, [dist_min + 2 * vac, 2 * vac, 2 * vac], True)
    for cls in (ab.Rebo2, ab.Rebo2Scr):
        calc = cls()
        e, f, st = calc.get_potential_energy(a), calc.get_forces(a).copy(), calc.get_stress(a).copy()
        b = a.copy()
        calc2 = cls()
        assert abs(e - calc2.get_potential_energy(b)) < 1e-10
        assert np.abs(f - calc2.get_forces(b)).max() < 1e-10
        assert np.abs(st - calc2.get_stress(b)).max() < 1e-10
        assert st.shape == (6,) and e < 0.0 and abs(f[0, 0] + f[1, 0]) < 1e-10
