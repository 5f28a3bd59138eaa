"""The CUDA kernels against the golden vectors of tests/golden/reference_executed.npz: outputs of the reference's own
Fortran kernels (executed through tests/fortran_subset.py where /root/reference exists; the vectors travel).  The
device is driven exactly as in the per-family parity tests (their _both helpers); only the thing compared with
changes: the reference's arithmetic instead of the oracle's.  1e-10 relative, neighbour lists entry by entry.
(The file sorts last on purpose: it was written after the round's GPU budget was spent.)"""
import numpy as np
import pytest

import reference_vectors as RV
from atomistica_b200 import native
from conftest import load_npz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', [c[0] for c in RV.cases('nl')])
def test_neighbor_lists(tag):
    a, _ = RV.atoms(tag)
    g = RV.outputs(tag)
    p = native.from_atoms(a)
    nl = native.Neighbors(200)
    nl.request_interaction_range(float(g['cutoff']))
    seed, last, nb, dc = nl.to_host(p)
    RV.check_list(seed, last, nb, dc, g, len(a))


@pytest.mark.parametrize('tag', [c[0] for c in RV.cases('eam')])
def test_eam(tag):
    import test_gpu_eam as G
    a, mask = RV.atoms(tag)
    g, _ = G._both(a, load_npz('cu_mishin1_setfl.npz'), mask=mask)
    RV.check_against(g, RV.outputs(tag), what=tag)


@pytest.mark.parametrize('case', RV.cases('bop'), ids=lambda c: c[0])
def test_bond_order_potentials(case):
    tag, _, kind, dbname, screened, _ = case
    a, mask = RV.atoms(tag)
    if screened:
        import test_gpu_bop_scr as G
        g, _ = G._both(kind, RV.parameter_set(dbname), a, mask=mask)
    else:
        import test_gpu_bop as G
        g, _, _ = G._both(kind, RV.parameter_set(dbname), a, mask=mask)
    e, f, w, epa, _, _, wpa, _ = g
    RV.check_against((e, f, w, epa, wpa), RV.outputs(tag), what=tag)


@pytest.mark.parametrize('case', RV.cases('rebo2'), ids=lambda c: c[0])
def test_rebo2(case):
    tag, _, _, _, screened, dihedral = case
    a, _ = RV.atoms(tag)
    if screened:
        import test_gpu_rebo2_scr as G
        g, _ = G._both(a, **(dict(with_dihedral=True) if dihedral else {}))
    else:
        import test_gpu_rebo2 as G
        g, _ = G._both(a, **(dict(dihedral=True) if dihedral else {}))
    e, f, w, epa, _, _, wpa, _ = g
    RV.check_against((e, f, w, epa, wpa), RV.outputs(tag), what=tag)


@pytest.mark.parametrize('case', RV.cases('juslin'), ids=lambda c: c[0])
def test_juslin(case):
    import test_gpu_juslin as G
    tag, _, _, dbname, screened, _ = case
    a, mask = RV.atoms(tag)
    g, _ = (G._both_scr if screened else G._both)(RV.parameter_set(dbname), a, mask=mask)
    e, f, w, epa, _, _, wpa, _ = g
    RV.check_against((e, f, w, epa, wpa), RV.outputs(tag), what=tag)
