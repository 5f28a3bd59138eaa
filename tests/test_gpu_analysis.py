"""Device-side list post-processing (csrc/atx_analysis.cu) against the array helpers of
atomistica_b200.analysis, which tests/test_analysis.py pins to a loop restatement of
src/python/c/analysis.c."""
import numpy as np
import pytest

from atomistica_b200 import analysis, native, structures as S
from conftest import load_npz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def system():
    d = load_npz('aC.npz')
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    p = native.from_atoms(a)
    nl = native.Neighbors(100)
    nl.request_interaction_range(3.0)
    i, j, dr, abs_dr = nl.get_neighbors(p, vec=True)
    return a, p, nl, i.astype(np.int32), j.astype(np.int32), dr, abs_dr


def test_coordination_numbers(system):
    a, p, nl, i, j, dr, abs_dr = system
    for rc in (1.85, 3.0):
        c = nl.get_coordination_numbers(p, rc)
        assert np.array_equal(c, np.bincount(i[abs_dr * abs_dr < rc * rc], minlength=len(a)))


def test_pair_distribution(system):
    a, p, nl, i, j, dr, abs_dr = system
    for nbins, cutoff in ((50, 3.0), (200, 2.5)):
        h, h2 = nl.pair_distribution(p, nbins, cutoff)
        g, g2 = analysis.pair_distribution(i, abs_dr, nbins, cutoff)
        assert np.abs(h - g).max() <= 1e-12 * np.abs(g).max()
        assert np.abs(h2 - g2).max() <= 1e-10 * max(np.abs(g2).max(), 1e-300)


def test_angle_distribution_and_moments(system):
    a, p, nl, i, j, dr, abs_dr = system
    h, h2 = nl.angle_distribution(p, 90, 1.85)
    g, g2 = analysis.angle_distribution(i, j, dr, 90, 1.85)
    assert np.abs(h - g).max() <= 1e-12 * np.abs(g).max()
    assert np.abs(h2 - g2).max() <= 1e-10 * np.abs(g2).max()
    for moment in (1, 2):
        m = nl.bond_angles(p, moment, 1.85)
        ref = analysis.bond_angles(moment, len(a), i, j, dr, 1.85)
        assert np.abs(m - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
