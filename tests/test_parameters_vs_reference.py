"""atomistica_b200.parameters against the reference's own src/python/atomistica/parameters.py, imported live
(it is pure Python without dependencies).  Runs where /root/reference exists (the build container); the
parameter values themselves are additionally exercised by the known-answer tests everywhere."""
import importlib.util
import os

import numpy as np
import pytest

from atomistica_b200 import parameters as P

REF = '/root/reference/src/python/atomistica/parameters.py'


@pytest.fixture(scope='module')
def ref():
    if not os.path.exists(REF):
        pytest.skip('the reference tree is not present')
    spec = importlib.util.spec_from_file_location('reference_parameters', REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _sets(m):
    # 'p' is the loop variable the reference's module leaves behind (an alias of its last set)
    return {n: getattr(m, n) for n in dir(m)
            if isinstance(getattr(m, n), dict) and 'el' in getattr(m, n) and not n.startswith('_') and n != 'p'}


def test_every_reference_parameter_set_exists_with_identical_values(ref):
    theirs, mine = _sets(ref), _sets(P)
    assert len(theirs) == 18
    for name, db in theirs.items():
        assert name in mine, name
        for key, val in db.items():
            if key == '__ref__':
                continue
            assert key in mine[name], (name, key)
            if key == 'el':
                assert list(mine[name][key]) == list(val), name
            else:
                got, want = np.asarray(mine[name][key], float), np.asarray(val, float)
                assert got.shape == want.shape, (name, key)
                assert np.array_equal(got, want), (name, key, got, want)      # bit-identical literals
        # nothing may be extra but citation text and fields spelled out with the value the Fortran type
        # declares as its default (tersoff_params.f90:66-80, juslin_params.f90:92): an extra key with another
        # value would change the potential
        for key in set(mine[name]) - set(db) - {'__ref__'}:
            default = 1 if key == 'm' else P.TERSOFF_FIELD_DEFAULTS.get(key)
            assert default is not None and all(v == default for v in mine[name][key]), (name, key)
