"""atomistica_b200.analysis against a loop-by-loop restatement of src/python/c/analysis.c."""
import numpy as np

import oracle
from atomistica_b200 import analysis, structures as S


def loops_pair_distribution(i, r, nbins, cutoff):
    # analysis.c:66-103
    h = np.zeros(nbins); h2 = np.zeros(nbins); tmp = np.zeros(nbins, dtype=np.int64)
    last_i = i[0]; nat = 1
    for p in range(len(i)):
        if last_i != i[p]:
            h += tmp; h2 += tmp * tmp; tmp[:] = 0; last_i = i[p]; nat += 1
        b = int(nbins * r[p] / cutoff)
        if 0 <= b < nbins:
            tmp[b] += 1
    for b in range(nbins):
        h[b] += tmp[b]; h2[b] += tmp[b] * tmp[b]
        r1, r2 = b * cutoff / nbins, (b + 1) * cutoff / nbins
        vol = 4 * np.pi / 3 * (r2 ** 3 - r1 ** 3)
        h[b] /= nat * vol; h2[b] /= nat * vol * vol; h2[b] -= h[b] * h[b]
    return h, h2


def loops_angles(i, r, cutoff):
    out = []
    last_i = i[0]; i_start = 0
    for p in range(len(i)):
        if last_i != i[p]:
            last_i = i[p]; i_start = p
        n = r[p] @ r[p]
        if n < cutoff * cutoff:
            p2 = i_start
            while p2 < len(i) and i[p2] == last_i:
                if p2 != p:
                    n2 = r[p2] @ r[p2]
                    if n2 < cutoff * cutoff:
                        out.append((last_i, np.arccos(r[p] @ r[p2] / np.sqrt(n * n2))))
                p2 += 1
    return out


def loops_angle_distribution(i, r, nbins, cutoff):
    # analysis.c:150-203
    ang = loops_angles(i, r, cutoff)
    atoms = sorted(set(i))
    cnt = {a: np.zeros(nbins) for a in atoms}
    for a, x in ang:
        cnt[a][int(nbins * x / np.pi) % nbins] += 1
    nangle = 1 + len(ang)
    vol = np.pi / nbins
    h = sum(cnt.values()) / (nangle * vol)
    h2 = sum(c * c for c in cnt.values()) / (nangle * vol * vol) - h * h
    return h, h2


def loops_bond_angles(moment, nat, i, r, cutoff):
    m = np.zeros(nat)
    acc, num = {}, {}
    for a, x in loops_angles(i, r, cutoff):
        acc[a] = acc.get(a, 0.0) + x ** moment
        num[a] = num.get(a, 0) + 1
    for a in acc:
        m[a] = acc[a] / num[a]
    return m


def _pairs(a, cutoff):
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 200)
    i, j, dc, _ = oracle.pairs(nl, len(a))
    dr = a.positions[i] - a.positions[j] + dc @ a.cell
    return i.astype(np.int32), j.astype(np.int32), dr, np.sqrt((dr * dr).sum(axis=1))


def test_pair_distribution(aC_small):
    i, j, dr, absdr = _pairs(aC_small, 4.0)
    for nbins, cutoff in ((50, 4.0), (17, 2.5), (8, 6.0)):
        h, h2 = analysis.pair_distribution(i, absdr, nbins, cutoff)
        rh, rh2 = loops_pair_distribution(i, absdr, nbins, cutoff)
        assert np.abs(h - rh).max() <= 1e-13 * np.abs(rh).max()
        assert np.abs(h2 - rh2).max() <= 1e-12 * np.abs(rh2).max()
    assert h.sum() > 0


def test_angle_distribution_and_bond_angles(aC_small):
    i, j, dr, absdr = _pairs(aC_small, 2.4)
    h, h2 = analysis.angle_distribution(i, j, dr, 36, 1.85)
    rh, rh2 = loops_angle_distribution(i, dr, 36, 1.85)
    assert np.abs(h - rh).max() <= 1e-13 * np.abs(rh).max()
    assert np.abs(h2 - rh2).max() <= 1e-12 * np.abs(rh2).max()
    for moment in (1, 2):
        m = analysis.bond_angles(moment, len(aC_small), i, j, dr, 1.85)
        rm = loops_bond_angles(moment, len(aC_small), i, dr, 1.85)
        assert np.abs(m - rm).max() <= 1e-13 * np.abs(rm).max()
    # amorphous carbon: mean bond angle between sp2 (120) and sp3 (109.5)
    mean = np.degrees(analysis.bond_angles(1, len(aC_small), i, j, dr, 1.85))
    assert 105 < mean[mean > 0].mean() < 122


def test_diamond_angles_are_tetrahedral():
    a = S.diamond('C', 3.566, (2, 2, 2))
    i, j, dr, absdr = _pairs(a, 2.0)
    m = analysis.bond_angles(1, len(a), i, j, dr, 1.8)
    assert np.allclose(np.degrees(m), 109.4712206, atol=1e-6)
    h, h2 = analysis.pair_distribution(i, absdr, 20, 2.0)
    assert np.count_nonzero(h) == 1     # one shell: the first neighbours


def test_empty_and_isolated_atoms():
    m = analysis.bond_angles(1, 5, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)), 2.0)
    assert (m == 0).all()
    # atom 3 has a single bond: no angle, stays 0
    i = np.array([0, 0, 3], dtype=np.int32); j = np.array([1, 2, 0], dtype=np.int32)
    dr = np.array([[1.0, 0, 0], [0, 1.0, 0], [1.0, 0, 0]])
    m = analysis.bond_angles(1, 4, i, j, dr, 2.0)
    assert abs(m[0] - np.pi / 2) < 1e-15 and m[3] == 0.0


def _reference_module():
    """the reference's own extension module (its unmodified src/python/c/analysis.c among the glue, built by
    atomistica_b200/seam2/build.py where /root/reference exists; the built module travels)"""
    import os
    import pytest
    from atomistica_b200.seam2 import build as s2build
    if os.path.isdir('/root/reference/src/python/c'):
        s2build.build()
    m = s2build.load()
    if m is None:
        pytest.skip('the reference-built _atomistica module is not present')
    return m


def test_against_the_reference_c_code(aC_small):
    """analysis.py (the checker of the device kernels in tests/test_gpu_analysis.py) against the REFERENCE's
    compiled analysis.c -- these three helpers are plain C on numpy arrays, so the reference itself runs here"""
    ref = _reference_module()
    i, j, dr, absdr = _pairs(aC_small, 2.6)
    dr = np.ascontiguousarray(dr)
    for nbins, cutoff in ((50, 2.6), (13, 2.0), (200, 2.6)):
        h, h2 = analysis.pair_distribution(i, absdr, nbins, cutoff)
        rh, rh2 = ref.pair_distribution(i, np.ascontiguousarray(absdr), nbins, cutoff)
        np.testing.assert_allclose(h, rh, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(h2, rh2, rtol=1e-10, atol=1e-13)
    for nbins, cutoff in ((36, 1.85), (90, 2.2)):
        h, h2 = analysis.angle_distribution(i, j, dr, nbins, cutoff)
        rh, rh2 = ref.angle_distribution(i, j, dr, nbins, cutoff)
        np.testing.assert_allclose(h, rh, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(h2, rh2, rtol=1e-10, atol=1e-13)
    for moment in (1, 2, 3):
        m = analysis.bond_angles(moment, len(aC_small), i, j, dr, 1.85)
        rm = ref.bond_angles(moment, len(aC_small), i, j, dr, 1.85)
        np.testing.assert_allclose(m, rm, rtol=1e-12, atol=1e-14)
