"""The oracle against the golden vectors of tests/golden/reference_executed.npz -- outputs of the reference's own
Fortran kernels executed through tests/fortran_subset.py.  Unlike tests/test_func_vs_reference.py this needs no
/root/reference, so it also runs on the GPU box; tests/test_zz_gpu_reference_vectors.py compares the CUDA kernels with
the same vectors."""
import numpy as np
import pytest

import oracle
import reference_vectors as RV
from atomistica_b200 import parameters as P
from conftest import load_npz

TOL = 1e-11


def _compare(o, g, per_bond=True):
    assert abs(o['epot'] - float(g['epot'])) <= TOL * abs(float(g['epot']))
    fs = max(1.0, np.abs(g['f']).max())
    ws = max(1.0, np.abs(g['wpot']).max(), abs(float(g['epot'])))
    assert np.abs(o['f'] - g['f']).max() <= TOL * fs
    assert np.abs(o['wpot'] - g['wpot']).max() <= TOL * ws
    assert np.abs(o['epot_per_at'] - g['epot_per_at']).max() <= TOL * max(1.0, np.abs(g['epot_per_at']).max())
    assert np.abs(o['wpot_per_at'] - g['wpot_per_at']).max() <= TOL * ws
    if per_bond:
        n = len(g['epot_per_bond'])
        assert n > 0
        assert np.abs(o['epot_per_bond'][:n] - g['epot_per_bond']).max() <= TOL * max(1.0, np.abs(g['epot_per_bond']).max())
        assert np.abs(o['f_per_bond'][:n] - g['f_per_bond']).max() <= TOL * fs
        assert np.abs(o['wpot_per_bond'][:n] - g['wpot_per_bond']).max() <= TOL * ws
        assert not np.any(o['epot_per_bond'][n:])


def test_the_vectors_are_complete():
    fams = [c[1] for c in RV.cases()]
    assert fams.count('nl') == 5 and fams.count('eam') == 2 and fams.count('bop') == 14
    assert fams.count('rebo2') == 8 and fams.count('juslin') == 4


@pytest.mark.parametrize('tag', [c[0] for c in RV.cases('nl')])
def test_neighbor_lists(tag):
    a, _ = RV.atoms(tag)
    g = RV.outputs(tag)
    nat = len(a)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, float(g['cutoff']), 200)
    n = len(g['neighbors'])
    assert np.array_equal(np.asarray(nl.seed), g['seed'])
    assert np.array_equal(np.asarray(nl.last)[:nat], g['last'][:nat])
    assert np.array_equal(np.asarray(nl.neighbors)[:n], g['neighbors'])
    real = g['neighbors'] != 0
    assert np.array_equal(np.asarray(nl.dc)[:n][real], g['dc'][real])


@pytest.mark.parametrize('tag', [c[0] for c in RV.cases('eam')])
def test_eam(tag):
    a, mask = RV.atoms(tag)
    eam = oracle.EAM(load_npz('cu_mishin1_setfl.npz'))
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff, 200)
    o = eam.energy_and_forces(a.positions, a.cell, nl, eam.eldb(a.symbols), mask=mask, per_at=True)
    _compare(o, RV.outputs(tag), per_bond=False)


@pytest.mark.parametrize('case', RV.cases('bop'), ids=lambda c: c[0])
def test_bond_order_potentials(case):
    tag, _, kind, dbname, screened, _ = case
    a, mask = RV.atoms(tag)
    okind = dict(Tersoff=oracle.TERSOFF, Kumagai=oracle.KUMAGAI, Brenner=oracle.BRENNER)[kind]
    db = (P.complete_scr if screened else P.complete)(kind, RV.parameter_set(dbname))
    nel = len(db['el'])
    if screened:
        cutoff = P.scr_cutoff(db)
    else:
        present = [db['el'].index(s) for s in set(a.symbols) if s in db['el']]
        cutoff = max(db['r2'][P.pair_index(i, j, nel)] for i in present for j in present)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 200)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(okind, db), a.positions, a.cell, nl, el, mask=mask, per_at=True,
                                     per_bond=True, scr=oracle.bop_scr_params(db) if screened else None)
    _compare(o, RV.outputs(tag))


@pytest.mark.parametrize('case', RV.cases('rebo2'), ids=lambda c: c[0])
def test_rebo2(case):
    tag, _, _, _, screened, dihedral = case
    a, _ = RV.atoms(tag)
    rb = (oracle.Rebo2Scr if screened else oracle.Rebo2)(with_dihedral=dihedral)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, rb.cutoff(a.symbols), 200)
    o = rb.energy_and_forces(a.positions, a.cell, nl, rb.ktyp(a.symbols), per_at=True, per_bond=True)
    _compare(o, RV.outputs(tag))


@pytest.mark.parametrize('case', RV.cases('juslin'), ids=lambda c: c[0])
def test_juslin(case):
    tag, _, _, dbname, screened, _ = case
    a, mask = RV.atoms(tag)
    db = (P.complete_juslin_scr if screened else P.complete_juslin)(RV.parameter_set(dbname))
    cutoff = P.juslin_scr_cutoff(db) if screened else max(db['r2'])
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 200)
    el = np.array([db['el'].index(s) + 1 if s in db['el'] else -1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.JUSLIN, db), a.positions, a.cell, nl, el, mask=mask,
                                     per_at=True, per_bond=True, scr=oracle.bop_scr_params(db) if screened else None)
    _compare(o, RV.outputs(tag))


def test_the_checkers_the_gpu_test_uses():
    """reference_vectors.check_against / check_list (used by tests/test_zz_gpu_reference_vectors.py on the GPU box) with
    the oracle standing in for the device: they accept matching results and reject a perturbed one"""
    tag = 'bop0_0'
    a, mask = RV.atoms(tag)
    db = P.complete('Tersoff', P.Tersoff_PRB_39_5566_Si_C)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.TERSOFF, db), a.positions, a.cell, nl, el, per_at=True)
    g = (o['epot'], o['f'], o['wpot'], o['epot_per_at'], o['wpot_per_at'])
    RV.check_against(g, RV.outputs(tag))
    bad = (o['epot'], o['f'] * (1 + 1e-8), o['wpot'], o['epot_per_at'], o['wpot_per_at'])
    with pytest.raises(AssertionError):
        RV.check_against(bad, RV.outputs(tag))
    a, _ = RV.atoms('nl2')
    gl = RV.outputs('nl2')
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, float(gl['cutoff']), 200)
    RV.check_list(nl.seed, nl.last, nl.neighbors, nl.dc, gl, len(a))
    nb = np.array(nl.neighbors).copy(); nb[3], nb[4] = nb[4], nb[3]
    with pytest.raises(AssertionError):
        RV.check_list(nl.seed, nl.last, nb, nl.dc, gl, len(a))


def test_gpu_file_dry_run_with_the_oracle_as_the_device(monkeypatch):
    """tests/test_zz_gpu_reference_vectors.py end to end without a GPU: the device-side helpers (_both of the per-family
    GPU test modules, native.from_atoms / Neighbors) are replaced by the oracle, so every case's routing, argument
    passing and unpacking is exercised here; on the GPU box the same functions run against the CUDA kernels"""
    import test_gpu_bop, test_gpu_bop_scr, test_gpu_eam, test_gpu_juslin, test_gpu_rebo2, test_gpu_rebo2_scr
    import test_zz_gpu_reference_vectors as Z
    from atomistica_b200 import native

    def as_g(o):
        return (o['epot'], o['f'], o['wpot'], o['epot_per_at'], None, None, o['wpot_per_at'], None)

    def bop(kind, db, a, mask=None, per_bond=False, avgn=100):
        db = P.complete(kind, db)
        okind = dict(Tersoff=oracle.TERSOFF, Kumagai=oracle.KUMAGAI, Brenner=oracle.BRENNER)[kind]
        idx = [db['el'].index(s) for s in set(a.symbols) if s in db['el']]
        nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2'][P.pair_index(i, j, len(db['el']))] for i in idx for j in idx), avgn)
        el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
        o = oracle.bop_energy_and_forces(oracle.bop_params(okind, db), a.positions, a.cell, nl, el, mask=mask, per_at=True)
        return as_g(o), o, nl

    def bop_scr(kind, db, a, mask=None, per_bond=False, avgn=1000):
        db = P.complete_scr(kind, db)
        okind = dict(Tersoff=oracle.TERSOFF, Kumagai=oracle.KUMAGAI, Brenner=oracle.BRENNER)[kind]
        nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, P.scr_cutoff(db), avgn)
        el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
        o = oracle.bop_energy_and_forces(oracle.bop_params(okind, db), a.positions, a.cell, nl, el, mask=mask, per_at=True,
                                         scr=oracle.bop_scr_params(db))
        return as_g(o), o

    def eam(a, setfl, mask=None, per_at=True):
        e = oracle.EAM(setfl)
        nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, e.cutoff, 200)
        o = e.energy_and_forces(a.positions, a.cell, nl, e.eldb(a.symbols), mask=mask, per_at=True)
        return (o['epot'], o['f'], o['wpot'], o['epot_per_at'], o['wpot_per_at']), o

    def rebo2(cls):
        def both(a, per_bond=False, **kw):
            if 'dihedral' in kw:
                kw['with_dihedral'] = kw.pop('dihedral')
            rb = cls(**kw)
            nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, rb.cutoff(a.symbols), 1000)
            o = rb.energy_and_forces(a.positions, a.cell, nl, rb.ktyp(a.symbols), per_at=True)
            return as_g(o), o
        return both

    def juslin(screened):
        def both(db, a, mask=None, per_bond=False):
            db = (P.complete_juslin_scr if screened else P.complete_juslin)(db)
            cutoff = P.juslin_scr_cutoff(db) if screened else max(db['r2'])
            nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 1000)
            el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
            o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.JUSLIN, db), a.positions, a.cell, nl, el, mask=mask,
                                             per_at=True, scr=oracle.bop_scr_params(db) if screened else None)
            return as_g(o), o
        return both

    class FakeNeighbors:
        def __init__(self, avgn):
            self.avgn = avgn

        def request_interaction_range(self, cutoff):
            self.cutoff = cutoff

        def to_host(self, a):
            nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, self.cutoff, self.avgn)
            return nl.seed, nl.last, nl.neighbors, nl.dc

    monkeypatch.setattr(test_gpu_bop, '_both', bop)
    monkeypatch.setattr(test_gpu_bop_scr, '_both', bop_scr)
    monkeypatch.setattr(test_gpu_eam, '_both', eam)
    monkeypatch.setattr(test_gpu_rebo2, '_both', rebo2(oracle.Rebo2))
    monkeypatch.setattr(test_gpu_rebo2_scr, '_both', rebo2(oracle.Rebo2Scr))
    monkeypatch.setattr(test_gpu_juslin, '_both', juslin(False))
    monkeypatch.setattr(test_gpu_juslin, '_both_scr', juslin(True))
    monkeypatch.setattr(native, 'from_atoms', lambda a: a)
    monkeypatch.setattr(native, 'Neighbors', FakeNeighbors)
    n = 0
    for c in RV.cases():
        fam = c[1]
        if fam == 'nl':
            Z.test_neighbor_lists(c[0])
        elif fam == 'eam':
            Z.test_eam(c[0])
        elif fam == 'bop':
            Z.test_bond_order_potentials(c)
        elif fam == 'rebo2':
            Z.test_rebo2(c)
        else:
            Z.test_juslin(c)
        n += 1
    assert n == 33
