import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope='session')
def cu_setfl():
    return load_npz('cu_mishin1_setfl.npz')


@pytest.fixture(scope='session')
def au_setfl():
    return load_npz('au_grochola_setfl.npz')


@pytest.fixture(scope='session')
def au_funcfl():
    return load_npz('au_u3_funcfl.npz')


@pytest.fixture(scope='session')
def aC():
    from atomistica_b200.structures import Atoms
    d = load_npz('aC.npz')
    return Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)


@pytest.fixture(scope='session')
def aC_small():
    from atomistica_b200.structures import Atoms
    d = load_npz('aC_small.npz')
    return Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
