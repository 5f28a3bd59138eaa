"""Host-side mirror of the reference interface: pure-Python pieces that need no GPU."""
import os

import numpy as np

from atomistica_b200 import io, parameters as P, structures as S
from atomistica_b200.elements import atomic_numbers

HERE = os.path.dirname(os.path.abspath(__file__))


def test_pair_index_matches_macros_inc():
    # PAIR_INDEX (macros.inc:123): symmetric, enumerates the upper triangle row by row
    for nel in (1, 2, 3):
        seen = {}
        for i in range(nel):
            for j in range(nel):
                k = P.pair_index(i, j, nel)
                assert k == P.pair_index(j, i, nel)
                seen.setdefault(k, set()).add((min(i, j), max(i, j)))
        assert sorted(seen) == list(range(nel * (nel + 1) // 2))
        assert all(len(v) == 1 for v in seen.values())


def test_complete_fills_fortran_defaults():
    db = P.complete('Tersoff', dict(el=['Si'], A=[1830.8], B=[471.18], r1=[2.7], r2=[3.0]))
    assert db['xi'] == [1.0] and db['mubo'] == [0.0] and db['m'] == [1] and db['n'] == [1.0]
    full = P.complete('Tersoff', None)
    assert full['el'] == ['C', 'Si'] and len(full['A']) == 3 and len(full['beta']) == 2
    scr = P.complete_scr('Tersoff', None)
    for k in P.SCR_KEYS:
        assert len(scr[k]) == 3
    assert scr['m'] == [3, 3, 3]
    # mubo = 1/dimer length, tersoff_params.f90:113
    assert np.allclose(scr['mubo'], [0.69103023078057590, 0.56580821386164815, 0.43569872294774004], rtol=1e-14)
    # default_bind_to_func.f90:44-66: Cmax = 3 -> C_dr_cut = 9/8, largest cutoff 6 A
    assert abs(P.scr_cutoff(scr) - np.sqrt(9.0 / 8.0) * 6.0) < 1e-12


def test_juslin_mirroring():
    db = P.complete_juslin(None)
    nel = 3
    for i in range(nel):
        for j in range(nel):
            a, b = j + i * nel, i + j * nel
            for key in P.JUSLIN_PAIR_KEYS:
                assert db[key][a] == db[key][b] or P.Juslin_JAP_98_123520_WCH['r0'][a] > 0 and \
                    P.Juslin_JAP_98_123520_WCH['r0'][b] > 0, (key, i, j)
    assert min(db['r0']) > 0 and min(db['S']) > 1.0
    assert len(db['alpha']) == 27 and len(db['omega']) == 27 and len(db['m']) == 27
    # the input dictionaries are not modified
    assert P.Juslin_JAP_98_123520_WCH['r0'][3] == -1.0


def test_structure_builders():
    for a, n, z, d in ((S.fcc('Cu', 3.615, (2, 2, 2)), 32, 12, 3.615 / np.sqrt(2)),
                       (S.bcc('W', 3.165, (2, 2, 2)), 16, 8, 3.165 * np.sqrt(3) / 2),
                       (S.sc('He', 1.0, (3, 3, 3)), 27, 6, 1.0),
                       (S.diamond('Si', 5.432, (2, 2, 2)), 64, 4, 5.432 * np.sqrt(3) / 4),
                       (S.b1(['Na', 'Cl'], 5.64, (2, 2, 2)), 64, 6, 5.64 / 2),
                       (S.b2(['W', 'C'], 2.7, (3, 3, 3)), 54, 8, 2.7 * np.sqrt(3) / 2),
                       (S.b3(['Si', 'C'], 4.36, (2, 2, 2)), 64, 4, 4.36 * np.sqrt(3) / 4)):
        assert len(a) == n
        L = np.diag(a.cell)
        dr = a.positions[None, :, :] - a.positions[:, None, :]
        dr -= np.round(dr / L) * L
        dist = np.sqrt((dr ** 2).sum(-1))
        np.fill_diagonal(dist, 1e9)
        assert np.allclose(dist.min(axis=1), d)
        assert np.all((np.abs(dist - d) < 1e-9).sum(axis=1) == z)
    a = S.b1(['Na', 'Cl'], 5.64, (1, 1, 1))
    dist = np.sqrt(((a.positions[4:, None] - a.positions[None, :4]) ** 2).sum(-1))
    assert dist.min() > 2.0      # the two sublattices do not coincide
    assert a.get_atomic_numbers().tolist() == [atomic_numbers['Na']] * 4 + [atomic_numbers['Cl']] * 4
    a.symbols[0] = 'K'           # cached numbers follow symbol assignment
    assert a.get_atomic_numbers()[0] == atomic_numbers['K']


def test_setfl_round_trip(tmp_path, cu_setfl):
    fn = str(tmp_path / 'cu.eam.alloy')
    io.write_setfl(fn, cu_setfl)
    back = io.read_setfl(fn)
    for k in ('F', 'rho', 'rphi'):
        assert np.array_equal(np.asarray(back[k]), np.asarray(cu_setfl[k]))
    assert back['nr'] == int(cu_setfl['nr']) and back['cutoff'] == float(cu_setfl['cutoff'])


def test_funcfl_reader(tmp_path, au_funcfl):
    fn = str(tmp_path / 'au.eam')
    with open(fn, 'w') as f:
        f.write(str(au_funcfl['comment']) + '\n')
        f.write('%d %.17g %.17g %s\n' % (int(au_funcfl['Znum']), float(au_funcfl['mass']), float(au_funcfl['a0']),
                                         str(au_funcfl['lattice'])))
        f.write('%d %.17g %d %.17g %.17g\n' % (int(au_funcfl['nF']), float(au_funcfl['dF']), int(au_funcfl['nr']),
                                               float(au_funcfl['dr']), float(au_funcfl['cutoff'])))
        for key in ('F', 'Z', 'rho'):
            v = np.asarray(au_funcfl[key])
            for i in range(0, len(v), 5):
                f.write(' '.join('%.17g' % x for x in v[i:i + 5]) + '\n')
    tab = io.read_funcfl(fn)
    assert tab['name'] == 'Au' and tab['nr'] == 500
    for key in ('F', 'Z', 'rho'):
        assert np.array_equal(tab[key], np.asarray(au_funcfl[key]))


def test_element_filters_without_gpu():
    # the filter logic of the pair styles is plain Python (filter_from_string, filter.f90:55-120)
    from atomistica_b200.native import _Pair
    el2Z = [atomic_numbers['Ar'], atomic_numbers['Kr']]      # ids 1, 2 by ascending Z
    assert _Pair._filter('*', el2Z) == 0b110
    assert _Pair._filter('Ar', el2Z) == 0b010
    assert _Pair._filter('Kr', el2Z) == 0b100
    assert _Pair._filter('Ar, Kr', el2Z) == 0b110
    assert _Pair._filter('Xe', el2Z) == 0          # known element that is not present: empty filter
    try:
        _Pair._filter('Qq', el2Z)
        assert False
    except RuntimeError as e:
        assert 'Unknown element' in str(e)
