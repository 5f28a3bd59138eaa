"""GPU neighbour list == oracle, entry by entry (stronger than "equal as sorted pair sets").

Cases follow the reference's tests/test_neighbor_list.py plus the geometry corner cases of
python_neighbors.f90 (n_cells floor of 3, multi-image stencils, clamped non-periodic cells,
atoms far outside the cell, triclinic cell)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S

pytestmark = pytest.mark.gpu


def _compare(atoms, cutoff, avgn=100):
    p = native.from_atoms(atoms)
    nl = native.Neighbors(avgn)
    nl.request_interaction_range(cutoff)
    seed, last, nb, dc = nl.to_host(p)
    ref = oracle.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, cutoff, avgn)
    nat = len(atoms)
    assert nl.info()['npairs'] == ref.npairs
    assert np.array_equal(seed[:nat + 1], ref.seed[:nat + 1])
    assert np.array_equal(last[:nat], ref.last[:nat])
    n = ref.npairs + nat
    assert np.array_equal(nb[:n], ref.neighbors[:n])
    assert np.array_equal(dc[:n], ref.dc[:n])
    return nl, p, ref


def test_si_diamond():
    _compare(S.diamond('Si', 5.432, (4, 4, 4)), 3.0)


def test_find_neighbor_and_counts():
    a = S.diamond('Si', 5.432, (3, 3, 3))
    nl, p, ref = _compare(a, 3.0)
    assert nl.get_number_of_all_neighbors(p) == 4 * len(a)
    assert np.all(nl.get_coordination_numbers(p, 3.0) == 4)
    seed, last, nb, dc = nl.to_host(p)
    j = int(nb[seed[0] - 1]) - 1                 # first neighbour of atom 0
    n1, n2 = nl.find_neighbor(p, 0, j)
    assert nb[n1] == j + 1 and seed[0] - 1 <= n1 < last[0]
    assert nb[n2] == 1 and seed[j] - 1 <= n2 < last[j]
    far = int(np.argmax(((a.positions - a.positions[0]) ** 2).sum(1)))
    assert nl.find_neighbor(p, 0, far) == (-2, -2)


def test_rebuild_after_motion_single_pass(cu_setfl):
    """from the second build on the counting pass parks the pairs in fixed-width rows (no second
    distance search); the list must stay entry-by-entry identical to the oracle, also when an atom
    outgrows its row (compressed region -> fallback to the two-pass build)"""
    a = S.fcc('Cu', 3.615, (6, 6, 6))
    a.rattle(0.05, seed=2)
    rc = float(cu_setfl['cutoff'])
    p = native.from_atoms(a)
    nl = native.Neighbors(400)
    nl.request_interaction_range(rc)
    rng = np.random.RandomState(5)
    for it in range(4):
        if it == 3:      # pull 60 atoms towards atom 0: longest list grows far beyond the previous one
            d = a.positions[1:61] - a.positions[0]
            a.positions[1:61] = a.positions[0] + 0.45 * d
        elif it:
            a.positions = a.positions + rng.normal(scale=0.05, size=a.positions.shape)
        p.coordinates[:, :] = a.positions
        p.I_changed_positions()
        seed, last, nb, dc = nl.to_host(p)
        ref = oracle.neighbor_list(a.positions, a.cell, a.pbc, rc, 400)
        nat = len(a)
        n = ref.npairs + nat
        assert nl.info()['npairs'] == ref.npairs
        assert np.array_equal(seed[:nat + 1], ref.seed[:nat + 1])
        assert np.array_equal(nb[:n], ref.neighbors[:n])
        assert np.array_equal(dc[:n], ref.dc[:n])
    assert nl.counters()[0] == 4


def test_si_diamond_rattled_skin():
    a = S.diamond('Si', 5.432, (5, 4, 3))
    a.rattle(0.1, seed=1)
    _compare(a, 3.5)


def test_fcc_cu(cu_setfl):
    a = S.fcc('Cu', 3.615, (6, 6, 6))
    a.rattle(0.05, seed=2)
    _compare(a, float(cu_setfl['cutoff']))
    _compare(a, float(cu_setfl['cutoff']) + 1.0, avgn=200)


def test_aC_triclinic(aC):
    # tests/test_neighbor_list.py:37-57
    nl, p, ref = _compare(aC, 5.0, avgn=200)
    i, j, dr, abs_dr = nl.get_neighbors(p, vec=True)
    s = np.linalg.solve(aC.cell.T, (aC.positions[i] - aC.positions[j]).T).T
    s -= np.round(s)
    dr_direct = s @ aC.cell
    assert np.all(np.abs(dr - dr_direct) < 1e-12)
    assert np.all(np.abs(abs_dr - np.sqrt((dr_direct ** 2).sum(1))) < 1e-12)


def test_aC_small(aC_small):
    _compare(aC_small, 2.0)
    _compare(aC_small, 4.0)


def test_tiny_cell_multiple_images():
    # 8 atoms, cell smaller than 3 cutoffs: n_cells = 3 with stencil half width 2
    a = S.diamond('Si', 5.432, (1, 1, 1))
    nl, p, ref = _compare(a, 3.3)
    assert nl.info()['n_cells'] == [3, 3, 3]
    assert nl.info()['stencil'] == [2, 2, 2]
    _compare(a, 6.0, avgn=400)   # self images


def test_pbc_variants():
    # tests/test_neighbor_list.py:59-101
    pos = [[0.1, 0.5, 0.5], [0.9, 0.5, 0.5]]
    for pbc, n in ((True, 2), (False, 0), ([False, False, True], 0), ([True, False, False], 2)):
        a = S.Atoms(['C', 'C'], pos, [1, 1, 1], pbc)
        nl, p, ref = _compare(a, 0.3)
        assert nl.info()['npairs'] == n


def test_partial_pbc_and_clamping(aC):
    a = aC.copy()
    a.pbc[:] = [True, False, False]
    _compare(a, 3.0)
    a.pbc[:] = False
    a.set_cell(a.cell * 0.9, scale_atoms=False)   # atoms outside a non-periodic cell are clamped
    _compare(a, 3.0)


def test_atoms_far_outside_cell(aC):
    # tests/test_neighbor_list.py:103-119
    a = aC.copy()
    a.positions[100] += 3 * a.cell[0]
    a.positions[7] += np.array([1, 3, -4]) @ a.cell
    a.positions[11] -= 7 * a.cell[2]
    _compare(a, 2.5)


def test_floating_point_edge_case():
    # tests/test_neighbor_list.py:152-174 (ChangeLog v0.10.2)
    pos = np.array([[-4.41173839e-52, 0.0, 0.0], [-4.41173839e-52, 2.26371743, 2.26371743],
                    [2.26371743, 0.0, 2.26371743], [2.26371743, 2.26371743, 0.0],
                    [1.13185872, 1.13185872, 1.13185872], [1.13185872, 3.39557615, 3.39557615],
                    [3.39557615, 1.13185872, 3.39557615], [3.39557615, 3.39557615, 1.13185872]])
    a = S.Atoms(['Si'] * 4 + ['C'] * 4, pos, [4.527434867899659] * 3, True)
    nl, p, ref = _compare(a, 3.0)
    assert (nl.get_coordination_numbers(p, 3.0) == 4).all()


def test_overflow_error():
    a = S.fcc('Cu', 3.615, (4, 4, 4))
    p = native.from_atoms(a)
    nl = native.Neighbors(10)
    nl.request_interaction_range(5.5)
    with pytest.raises(RuntimeError, match='Neighbor list overflow'):
        nl.update(p)
    with pytest.raises(RuntimeError):
        oracle.neighbor_list(a.positions, a.cell, a.pbc, 5.5, 10)


def test_empty_and_single():
    a = S.Atoms(['Si'], [[0.3, 0.3, 0.3]], [10, 10, 10], True)
    nl, p, ref = _compare(a, 3.0)
    assert nl.info()['npairs'] == 0


@pytest.mark.parametrize('n', [20, 40])
def test_large_property(n):
    # size-independent properties at sizes the oracle does not need to run: every atom of a perfect
    # diamond lattice has exactly 4 neighbours at a*sqrt(3)/4, and the list is symmetric
    a = S.diamond('Si', 5.432, (n, n, n))
    p = native.from_atoms(a)
    nl = native.neighbor_list(p, 3.0)
    assert nl.info()['npairs'] == 4 * len(a)
    assert nl.info()['nebmax'] == 4


# ---- well-filled cells: the warp-per-cell kernel (>= 6 atoms per cell, stencil of +-1 cell) -------------

def _gas(n, box, seed, pbc=True, cell=None):
    rng = np.random.RandomState(seed)
    pos = rng.rand(n, 3) * np.asarray(box)
    return S.Atoms(['Cu'] * n, pos, np.diag(box) if cell is None else cell, pbc)


@pytest.mark.parametrize('pbc', [True, False, [True, False, True], [False, True, False]])
def test_dense_cells_pbc_variants(pbc):
    a = _gas(400, [12.4, 9.3, 9.8], 11, pbc)           # 4 x 3 x 3 cells at cutoff 3.0: 11 atoms per cell
    nl, p, ref = _compare(a, 3.0, avgn=200)
    assert nl.info()['n_cells'] == [4, 3, 3] and nl.info()['stencil'] == [1, 1, 1]
    # moved atoms, second build (fixed-width rows filled by the counting pass)
    a.positions += np.random.RandomState(3).normal(scale=0.1, size=a.positions.shape)
    p.coordinates[:, :] = a.positions
    p.I_changed_positions()
    seed, last, nb, dc = nl.to_host(p)
    ref = oracle.neighbor_list(a.positions, a.cell, a.pbc, 3.0, 200)
    n = ref.npairs + len(a)
    assert np.array_equal(nb[:n], ref.neighbors[:n]) and np.array_equal(dc[:n], ref.dc[:n])


def test_dense_cells_more_than_a_warp_per_cell():
    a = _gas(1300, [9.1, 9.2, 9.3], 5)                  # 3 x 3 x 3 cells: 48 atoms per cell, 2 groups of lanes
    nl, p, ref = _compare(a, 3.0, avgn=400)
    assert nl.info()['n_cells'] == [3, 3, 3] and nl.info()['stencil'] == [1, 1, 1]


def test_dense_cells_triclinic_and_band():
    cell = np.array([[12.0, 0.0, 0.0], [2.5, 9.0, 0.0], [1.0, -1.5, 8.0]])
    rng = np.random.RandomState(8)
    pos = rng.rand(260, 3) @ cell
    a = S.Atoms(['Cu'] * 260, pos, cell, True)
    # pairs placed within 1e-9 of the cutoff sphere (inside and outside): the band of the single-precision
    # test, decided by the exact predicate
    rc = 2.9
    for k in range(40):
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        a.positions[2 * k + 1] = a.positions[2 * k] + u * (rc + (1e-9 if k % 2 else -1e-9) * (1 + k))
    _compare(a, rc, avgn=200)


def test_dense_cells_match_the_thread_per_atom_kernel(monkeypatch):
    a = _gas(4000, [30.0, 28.0, 31.0], 9)
    p = native.from_atoms(a)
    out = []
    for coop in ('1', '0'):
        monkeypatch.setenv('ATX_NL_COOP', coop)
        nl = native.Neighbors(200)
        nl.request_interaction_range(3.2)
        out.append(nl.to_host(p))
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(x, y)
