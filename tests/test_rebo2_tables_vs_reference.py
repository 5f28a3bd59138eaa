"""The REBO2 default tables (Brenner 2002, Tables 4-9) of atomistica_b200.rebo2_tables and of the oracle against
the reference's own rebo2_default_tables.f90, which is EXECUTED here: the file consists of array assignments
and DO loops only, so a thirty-line translator turns each subroutine into Python.  The 36 atomisation energies
pin the entries molecules reach; this pins every entry, bit for bit.  Runs where /root/reference exists."""
import os
import re

import numpy as np
import pytest

import oracle
from atomistica_b200 import rebo2_tables

SRC = '/root/reference/src/potentials/bop/rebo2/rebo2_default_tables.f90'


def _index(expr, one_based_upper):
    """a Fortran subscript list -> numpy subscripts (lower bounds are 0 in this file; a:b is inclusive)"""
    out = []
    for sub in expr.split(','):
        sub = sub.strip()
        if sub == ':':
            out.append(':')
        elif ':' in sub:
            a, b = sub.split(':')
            out.append('%s:(%s)+1' % (a.strip(), b.strip()))
        else:
            out.append(sub)
    return ', '.join(out)


def run_fortran_subroutines(text):
    """{subroutine name: {array name: ndarray}}"""
    results, arrays, body, depth, name, skip = {}, None, None, 0, None, False
    for raw in text.splitlines():
        line = raw.split('!')[0].rstrip()
        if line.startswith('#ifdef ZERO_TABLES'):
            skip = True
            continue
        if line.startswith('#endif'):
            skip = False
            continue
        if skip or not line.strip():
            continue
        stmt = line.strip()
        m = re.match(r'subroutine (\w+)\(', stmt)
        if m:
            name, arrays, body, depth = m.group(1), {}, [], 0
            continue
        if name is None:
            continue
        if stmt.startswith('endsubroutine'):
            env = {'np': np}
            for a, shape in arrays.items():
                env[a] = np.zeros(shape)
            exec('\n'.join(body) or 'pass', env)
            results[name] = {a: env[a] for a in arrays}
            name = None
            continue
        m = re.match(r'real\(DP\), intent\(out\)\s*::\s*(\w+)\(([^)]*)\)', stmt)
        if m:
            dims = [d.strip().split(':') for d in m.group(2).split(',')]
            assert all(d[0] == '0' for d in dims), stmt
            arrays[m.group(1)] = tuple(int(d[1]) + 1 for d in dims)
            continue
        if stmt in ('implicit none',) or re.match(r'(real\(DP\)|integer)\s*::', stmt):
            continue
        m = re.match(r'do (\w+) = (.+), (.+)$', stmt)
        if m:
            body.append('    ' * depth + 'for %s in range(%s, (%s)+1):' % m.groups())
            depth += 1
            continue
        if stmt == 'enddo':
            depth -= 1
            continue
        # assignment: literals lose their kind suffix, array references become numpy subscripts
        py = re.sub(r'(\d)_DP\b', r'\1', stmt)
        py = re.sub(r'\b(%s)\(([^()]*)\)' % '|'.join(arrays), lambda r: '%s[%s]' % (r.group(1), _index(r.group(2), True)), py)
        assert '=' in py and '_DP' not in py, stmt
        body.append('    ' * depth + py)
    return results


@pytest.fixture(scope='module')
def reference_tables():
    if not os.path.exists(SRC):
        pytest.skip('the reference tree is not present')
    r = run_fortran_subroutines(open(SRC).read())
    assert sorted(r) == ['rebo2_default_Fcc_table', 'rebo2_default_Fch_table', 'rebo2_default_Fhh_table',
                         'rebo2_default_Pcc_table', 'rebo2_default_Pch_table', 'rebo2_default_Tcc_table']
    f = r['rebo2_default_Fcc_table']
    return dict(Fcc=f['F'], dFdi=f['dFdi'], dFdj=f['dFdj'], dFdk=f['dFdk'], Fch=r['rebo2_default_Fch_table']['F'],
                Fhh=r['rebo2_default_Fhh_table']['F'], Pcc=r['rebo2_default_Pcc_table']['P'],
                Pch=r['rebo2_default_Pch_table']['P'], Tcc=r['rebo2_default_Tcc_table']['T'])


def test_translator_sees_the_tables(reference_tables):
    t = reference_tables
    assert t['Fcc'].shape == (5, 5, 10) and t['Pcc'].shape == (6, 6)
    assert t['Fcc'][1, 1, 0] == 0.105 and t['Fcc'][1, 0, 0] == 0.04338699 == t['Fcc'][0, 1, 0]   # symmetrised
    assert t['Fcc'][2, 2, 5] == 0.03970587 + 3 * (0.0 - 0.03970587) / 6                 # the DO loop ran
    assert t['Tcc'][2, 2, 8] == -0.00809675 and t['Tcc'][2, 2, 9] == 0.0                 # inclusive upper bound
    assert sum(int(np.count_nonzero(v)) for v in t.values()) > 150


@pytest.mark.parametrize('which', ['product', 'oracle'])
def test_default_tables_equal_the_reference_bit_for_bit(reference_tables, which):
    mine = rebo2_tables.default_tables() if which == 'product' else oracle.rebo2_default_tables()
    for key, ref in reference_tables.items():
        got = np.asarray(mine[key])
        assert got.shape == ref.shape, key
        assert np.array_equal(got, ref), (key, np.argwhere(got != ref)[:5])


TYPE_SRC = '/root/reference/src/potentials/bop/rebo2/rebo2_type.f90'


def _type_defaults(screening=False):
    """component defaults of rebo2_type.f90 for a build without / with SCREENING: scalars, array constructors; a
    literal without kind suffix is default real (single precision) promoted to double, as the compiler does"""
    if not os.path.exists(TYPE_SRC):
        pytest.skip('the reference tree is not present')
    text, out, active = [], {}, [True]
    for raw in open(TYPE_SRC).read().splitlines():
        if raw.startswith('#ifndef'):              # no macro is defined in the plain build
            active.append(True)
        elif raw.startswith('#if'):
            active.append(screening and raw.startswith('#ifdef SCREENING'))
        elif raw.startswith('#else'):
            active[-1] = not active[-1]
        elif raw.startswith('#endif'):
            active.pop()
        elif all(active):
            text.append(raw.split('!')[0].rstrip())
    joined = re.sub(r'&\s*\n\s*', ' ', '\n'.join(text))

    def literal(tok):
        tok = tok.strip()
        m = re.fullmatch(r'(-?[\d.]+(?:[eEdD][-+]?\d+)?)(_DP)?(?:/(\d+))?', tok)
        assert m, tok
        v = float(m.group(1).replace('d', 'e').replace('D', 'e'))
        if not m.group(2) and ('.' in m.group(1)):
            v = float(np.float32(v))
        return v / int(m.group(3)) if m.group(3) else v

    for m in re.finditer(r'real\(DP\)\s*::\s*(\w+)\s*=\s*(-?[\d.]+(?:[eEdD][-+]?\d+)?(?:_DP)?)\s*$', joined, re.M):
        out[m.group(1)] = literal(m.group(2))
    for m in re.finditer(r'real\(DP\)\s*::\s*(\w+)\((\d+)\)\s*=\s*\(/(.*?)/\)', joined):
        out[m.group(1)] = np.array([literal(t) for t in m.group(3).split(',')])
    m = re.search(r'SPGH\(6,3\)\s*=\s*reshape\(\s*\(/(.*?)/\)', joined, re.S)
    out['SPGH'] = np.array([literal(t) for t in m.group(1).split(',')])
    m = re.search(r'integer\s*::\s*IGH\(25\)\s*=\s*\(/(.*?)/\)', joined, re.S)
    out['IGH'] = [int(t) for t in m.group(1).split(',')]
    return out


def test_scalar_defaults_and_spline_nodes_equal_the_reference():
    ref = _type_defaults()
    assert len(ref) >= 35
    for key, val in rebo2_tables.DEFAULTS.items():
        if key == 'dihedral':
            continue
        assert key in ref, key
        assert ref[key] == val, (key, ref[key], val)
    for mine, key in ((rebo2_tables.G_THETA, 'cc_g_theta'), (rebo2_tables.G_G1, 'cc_g_g1'), (rebo2_tables.G_DG1, 'cc_g_dg1'),
                      (rebo2_tables.G_D2G1, 'cc_g_d2g1'), (rebo2_tables.G_G2, 'cc_g_g2'), (rebo2_tables.SPGH, 'SPGH')):
        assert np.array_equal(np.asarray(mine, float), ref[key]), key
    assert list(rebo2_tables.IGH) == ref['IGH']
    # the single-precision reading matters: -0.01 as default real is not -0.01d0
    assert ref['cc_g_g1'][0] != -0.01 and abs(ref['cc_g_g1'][0] + 0.01) < 1e-9
    # the oracle keeps its own copy of the same constants
    for key, val in oracle.REBO2_DEFAULTS.items():
        if key != 'with_dihedral':
            assert ref[key] == val, key
    for mine, key in ((oracle.CC_G_THETA, 'cc_g_theta'), (oracle.CC_G_G1, 'cc_g_g1'), (oracle.CC_G_DG1, 'cc_g_dg1'),
                      (oracle.CC_G_D2G1, 'cc_g_d2g1'), (oracle.CC_G_G2, 'cc_g_g2'), (oracle.SPGH, 'SPGH')):
        assert np.array_equal(np.asarray(mine, float), ref[key]), key
    assert list(oracle.IGH) == ref['IGH']


def test_screened_defaults_equal_the_reference():
    ref = _type_defaults(screening=True)
    for table in (rebo2_tables.SCR_DEFAULTS, oracle.REBO2_SCR_DEFAULTS):
        assert set(rebo2_tables.SCR_KEYS) <= set(table)
        for key, val in table.items():
            assert ref[key] == val, (key, ref[key], val)
    assert ref['cc_in_r1'] == 1.95 and _type_defaults()['cc_in_r1'] == 1.70
