"""Seam 2 of the boundary: the reference's OWN `_atomistica` extension (its unmodified C glue from
/root/reference/src/python/c, built without Fortran by atomistica_b200/seam2/build.py) on top of the
B200 library.  The module is driven the way src/python/atomistica/native.py:36-63 and
aseinterface.py:227-289, 382 drive it; the GPU cases reproduce the reference's tests/test_tersoff.py,
tests/test_mask.py and tests/test_neighbor_list.py through it and compare with the oracle.
"""
import os

import numpy as np
import pytest

from atomistica_b200 import io as aio, structures as S
from atomistica_b200.elements import atomic_numbers
from atomistica_b200.seam2 import build as s2build
from conftest import ROOT, load_npz


@pytest.fixture(scope='module')
def mod():
    if os.path.isdir('/root/reference/src/python/c'):
        s2build.build()
    m = s2build.load()
    if m is None:
        pytest.skip('the reference-built _atomistica module is not present (built where /root/reference exists)')
    m.startup()
    return m


def test_module_exports(mod):
    """the reference's extension types, created by its own atomisticamodule.c from potential_classes[]"""
    for name in ('Particles', 'Neighbors', 'Tersoff', 'TersoffScr', 'Kumagai', 'KumagaiScr', 'Brenner', 'BrennerScr',
                 'TabulatedAlloyEAM', 'TabulatedEAM', 'Juslin', 'JuslinScr', 'Rebo2', 'Rebo2Scr', 'LJCut', 'Harmonic', 'DoubleHarmonic', 'BornMayer', 'r6',
                 'startup', 'shutdown', 'set_logfile', 'pair_distribution'):
        assert hasattr(mod, name), name
    assert mod.Tersoff.__name__ == 'Tersoff'


def test_shim_exports_every_symbol_the_glue_needs():
    """nm: no undefined symbol of the module is left for a Fortran runtime to supply"""
    import subprocess
    if not s2build.available():
        pytest.skip('module not built')
    out = subprocess.run(['nm', '-D', '--undefined-only', s2build.module_path()], capture_output=True, text=True).stdout
    und = [l.split()[-1] for l in out.splitlines() if l.strip()]
    fortran_side = [u for u in und if u.startswith(('f_', 'python_', 'data_', 'c_p', 'c_e')) or u.endswith('_by_name')]
    assert not fortran_side, fortran_side
    atx = [u for u in und if u.startswith('atx_')]
    assert atx, 'the module must resolve the compute entry points from libatomistica_b200.so'


def test_no_cpu_fallback(mod):
    import ctypes
    try:
        ctypes.CDLL('libcuda.so.1')
        has_gpu = os.path.exists('/dev/nvidia0')
    except OSError:
        has_gpu = False
    if has_gpu:
        pytest.skip('a GPU is present')
    p = mod.Particles()
    with pytest.raises(RuntimeError, match='No CUDA device'):
        p.allocate(8)
    with pytest.raises(RuntimeError, match='No CUDA device'):
        mod.Tersoff()


# ---------------------------------------------------------------------------------------------
# GPU: the reference's tests through its own module
# ---------------------------------------------------------------------------------------------

def _particles(mod, atoms):
    """atomistica/native.py:36-53 (from_atoms)"""
    p = mod.Particles()
    p.allocate(len(atoms))
    p.set_cell(atoms.cell, atoms.pbc)
    Z = p.Z
    Z[:] = [atomic_numbers[s] for s in atoms.symbols]
    r = p.coordinates
    r[:, :] = atoms.positions
    p.I_changed_positions()
    p.update_elements()
    return p


def _calc(mod, pot, atoms, avgn=100, **kw):
    p = _particles(mod, atoms)
    nl = mod.Neighbors(avgn)
    pot.bind_to(p, nl)
    return (p, nl) + tuple(pot.energy_and_forces(p, nl, **kw))


@pytest.mark.gpu
def test_tersoff_smoke_and_oracle(mod):
    """tests/test_tersoff.py:36-49 (bulk Si: finite energy, forces shape) + parity with the oracle"""
    import oracle
    from atomistica_b200 import parameters as P
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.05, seed=7)
    p, nl, e, f, w = _calc(mod, mod.Tersoff(), a)
    assert np.isfinite(e) and f.shape == (len(a), 3) and w.shape == (3, 3)
    db = P.complete('Tersoff', None)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.TERSOFF, db), a.positions, a.cell, onl, el)
    assert abs(e - o['epot']) <= 1e-10 * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())
    assert np.abs(w - o['wpot']).max() <= 1e-10 * max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))


@pytest.mark.gpu
def test_keyword_parameters_reach_the_device(mod):
    """parameters written through the ptrdict registry (potential.c:114-116, tersoff_registry.f90) are the ones
    the kernel uses: the Erhart-Albe set by keyword == the built-in Brenner default"""
    from atomistica_b200 import parameters as P
    a = S.b3(['Si', 'C'], 4.36, (2, 2, 2))
    a.rattle(0.05, seed=3)
    e0 = _calc(mod, mod.Brenner(), a)[2]
    kw = {k: v for k, v in P.Erhart_PRB_71_035211_SiC.items() if not k.startswith('__')}
    pot = mod.Brenner(**kw)
    e1 = _calc(mod, pot, a)[2]
    assert abs(e0 - e1) <= 1e-12 * abs(e0)
    assert list(pot.el) == ['C', 'Si'] and abs(pot.D0[1] - 4.36) < 1e-15      # potential_getattro
    kw['D0'] = [6.0, 4.0, 3.24]
    e2 = _calc(mod, mod.Brenner(**kw), a)[2]
    assert abs(e2 - e0) > 1e-3


@pytest.mark.gpu
def test_mask_decomposition(mod, tmp_path):
    """tests/test_mask.py:35-81: E, f, wpot with mask + complement == unmasked (1e-6), Tersoff / TersoffScr on
    the a-C fixture and TabulatedAlloyEAM on fcc Au"""
    d = load_npz('aC.npz')
    aC = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    fn = str(tmp_path / 'Au.eam.alloy')
    aio.write_setfl(fn, load_npz('au_grochola_setfl.npz'))
    au = S.fcc('Au', 4.07, (2, 2, 2))
    rng = np.random.RandomState(5)
    for make, atoms, avgn in ((lambda: mod.Tersoff(), aC, 100), (lambda: mod.TersoffScr(), aC, 1000),
                              (lambda: mod.TabulatedAlloyEAM(fn=fn), au, 100)):
        pot = make()
        p = _particles(mod, atoms)
        nl = mod.Neighbors(avgn)
        pot.bind_to(p, nl)
        e, f, w = pot.energy_and_forces(p, nl)
        mask = (rng.randint(0, len(atoms), size=len(atoms)) < len(atoms) / 2).astype(np.int32)
        e1, f1, w1 = pot.energy_and_forces(p, nl, mask=mask)
        e2, f2, w2 = pot.energy_and_forces(p, nl, mask=(1 - mask).astype(np.int32))
        assert abs(e - e1 - e2) < 1e-6
        assert np.abs(f - f1 - f2).max() < 1e-6
        assert np.abs(w - w1 - w2).max() < 1e-6


@pytest.mark.gpu
def test_neighbor_list(mod):
    """tests/test_neighbor_list.py:37-57: a-C fixture, cutoff 5.0 -- distances of the list equal the minimum
    image distances (1e-12) and the pair count equals a brute-force count"""
    d = load_npz('aC.npz')
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    p = _particles(mod, a)
    nl = mod.Neighbors(100)
    nl.request_interaction_range(5.0)
    i, j, abs_dr_no_vec = nl.get_neighbors(p)
    i, j, dr, abs_dr = nl.get_neighbors(p, vec=True)
    assert np.all(np.abs(abs_dr - abs_dr_no_vec) < 1e-12)
    r = a.positions
    dr_direct = r[i] - r[j]
    s = np.linalg.solve(a.cell.T, dr_direct.T).T
    s -= np.round(s)
    mic = s @ a.cell
    assert np.all(np.abs(np.sqrt((mic * mic).sum(axis=1)) - abs_dr) < 1e-12)
    assert np.all(np.abs(mic - dr) < 1e-12)
    # brute-force pair count on a sample of atoms
    for k in (0, 17, 2000, 4000):
        dk = r - r[k]
        sk = np.linalg.solve(a.cell.T, dk.T).T
        sk -= np.round(sk)
        dist = np.sqrt(((sk @ a.cell) ** 2).sum(axis=1))
        assert (i == k).sum() == ((dist < 5.0).sum() - 1)
    # coordination numbers and per-atom access (neighbors.c:get_neighbors(i=...), coordination)
    j17, r17 = nl.get_neighbors(p, 17)
    assert len(j17) == (i == 17).sum() and np.all(np.abs(np.sort(r17) - np.sort(abs_dr[i == 17])) < 1e-12)


@pytest.mark.gpu
def test_per_atom_and_per_bond_outputs(mod):
    import oracle
    from atomistica_b200 import parameters as P
    d = load_npz('aC_small.npz')
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    pot = mod.Tersoff()
    p = _particles(mod, a)
    nl = mod.Neighbors(100)
    pot.bind_to(p, nl)
    e, f, w, epa, epb, fpb, wpa, wpb = pot.energy_and_forces(p, nl, epot_per_at=True, epot_per_bond=True,
                                                             f_per_bond=True, wpot_per_at=True, wpot_per_bond=True)
    assert abs(epa.sum() - e) <= 1e-10 * abs(e)
    assert abs(epb.sum() - e) <= 1e-10 * abs(e)
    assert np.abs(wpa.sum(axis=0) - w).max() <= 1e-9 * max(1.0, abs(e))
    db = P.complete('Tersoff', None)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.TERSOFF, db), a.positions, a.cell, onl, el, per_at=True)
    assert np.abs(epa - o['epot_per_at']).max() <= 1e-10 * np.abs(o['epot_per_at']).max()


@pytest.mark.gpu
def test_rebo2_and_pair_classes(mod):
    """Rebo2 / Rebo2Scr (default parameter block generated at build time) and LJCut through the module == oracle"""
    import oracle
    d = load_npz('aC_small.npz')
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    for cls, ocls, avgn in ((mod.Rebo2, oracle.Rebo2, 100), (mod.Rebo2Scr, oracle.Rebo2Scr, 1000)):
        p, nl, e, f, w = _calc(mod, cls(), a, avgn=avgn)
        rb = ocls()
        onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, rb.cutoff(a.symbols), avgn)
        o = rb.energy_and_forces(a.positions, a.cell, onl, rb.ktyp(a.symbols))
        assert abs(e - o['epot']) <= 1e-10 * abs(o['epot'])
        assert np.abs(f - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())
    with pytest.raises(RuntimeError):
        pot = mod.Rebo2()
        p = _particles(mod, a)
        nl = mod.Neighbors(100)
        pot.bind_to(p, nl)
        pot.energy_and_forces(p, nl, mask=np.ones(len(a), dtype=np.int32))
    au = S.fcc('Au', 4.07, (3, 3, 3))
    au.rattle(0.05, seed=4)
    p, nl, e, f, w = _calc(mod, mod.LJCut(el1='Au', el2='Au', epsilon=1.0, sigma=2.6, cutoff=6.0), au, avgn=200)
    assert np.isfinite(e) and e < 0.0 and np.abs(f.sum(axis=0)).max() < 1e-9
    from atomistica_b200 import native
    pn = native.from_atoms(au)
    nn = native.Neighbors(200)
    lj = native.LJCut(el1='Au', el2='Au', epsilon=1.0, sigma=2.6, cutoff=6.0)
    lj.bind_to(pn, nn)
    e2, f2 = lj.energy_and_forces(pn, nn)[:2]
    assert abs(e - e2) <= 1e-12 * abs(e2) and np.abs(f - f2).max() <= 1e-12 * max(1.0, np.abs(f2).max())


@pytest.mark.gpu
def test_juslin_classes_and_funcfl(mod, tmp_path):
    """Juslin, JuslinScr (database through the ptrdict registry, mirrored by init like BIND_TO_FUNC) and
    TabulatedEAM (funcfl file read by the shim) through the reference's module == the Python mirror on
    the same device kernels == the oracle (tests/test_gpu_juslin.py, tests/test_gpu_eam.py)"""
    from atomistica_b200 import native, parameters as P
    a = S.b1(['W', 'C'], 4.38, (3, 3, 3))
    rng = np.random.RandomState(5)
    for i in rng.choice(len(a), 20, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=9)

    def mirror(pot, avgn):
        pn = native.from_atoms(a)
        nn = native.Neighbors(avgn)
        pot.bind_to(pn, nn)
        return pot.energy_and_forces(pn, nn)[:3]

    p, nl, e, f, w = _calc(mod, mod.Juslin(), a, avgn=200)
    e0, f0, w0 = mirror(native.Juslin(), 200)
    assert abs(e - e0) <= 1e-12 * abs(e0) and np.abs(f - f0).max() <= 1e-12 * max(1.0, np.abs(f0).max())
    assert np.abs(np.asarray(w).reshape(3, 3) - w0).max() <= 1e-10 * max(1.0, np.abs(w0).max())
    # screened class: default database, then wider outer / bond-order cutoffs by keyword
    p, nl, e, f, w = _calc(mod, mod.JuslinScr(), a, avgn=1000)
    e0, f0, w0 = mirror(native.JuslinScr(), 1000)
    assert abs(e - e0) <= 1e-12 * abs(e0) and np.abs(f - f0).max() <= 1e-12 * max(1.0, np.abs(f0).max())
    raw = {k: list(v) for k, v in P.Juslin_WCH__Scr_fortran_default.items() if not k.startswith('__')}
    for k in range(9):
        if raw['r2'][k] > 0:
            raw['or1'][k], raw['or2'][k] = raw['r2'][k] * 1.05, raw['r2'][k] * 1.45
            raw['bor1'][k], raw['bor2'][k] = raw['r2'][k] * 1.0, raw['r2'][k] * 1.35
    kw = {k: raw[k] for k in ('or1', 'or2', 'bor1', 'bor2')}
    p, nl, e1, f1, w1 = _calc(mod, mod.JuslinScr(**kw), a, avgn=1000)
    e2, f2, w2 = mirror(native.JuslinScr(P.complete_juslin_scr(raw)), 1000)
    assert abs(e1 - e) > 1e-6
    assert abs(e1 - e2) <= 1e-12 * abs(e2) and np.abs(f1 - f2).max() <= 1e-12 * max(1.0, np.abs(f2).max())

    # funcfl
    tab = load_npz('au_u3_funcfl.npz')
    fn = str(tmp_path / 'Au_u3.eam')
    with open(fn, 'w') as fh:
        fh.write('Au funcfl\n%d %.6f %.6f fcc\n' % (79, 196.97, 4.08))
        fh.write('%d %.17g %d %.17g %.17g\n' % (int(tab['nF']), float(tab['dF']), int(tab['nr']), float(tab['dr']),
                                                 float(tab['cutoff'])))
        for key in ('F', 'Z', 'rho'):
            v = np.asarray(tab[key], dtype=float)
            for k in range(0, len(v), 5):
                fh.write(' '.join('%.17g' % x for x in v[k:k + 5]) + '\n')
    au = S.fcc('Au', 4.08, (3, 3, 3))
    au.rattle(0.1, seed=41)
    p, nl, e, f, w = _calc(mod, mod.TabulatedEAM(fn=fn), au, avgn=200)
    pn = native.from_atoms(au)
    nn = native.Neighbors(200)
    pot = native.TabulatedEAM(funcfl=tab)
    pot.bind_to(pn, nn)
    e0, f0 = pot.energy_and_forces(pn, nn)[:2]
    assert abs(e - e0) <= 1e-12 * abs(e0) and np.abs(f - f0).max() <= 1e-12 * max(1.0, np.abs(f0).max())
    with pytest.raises(RuntimeError):
        mod.TabulatedEAM(fn=str(tmp_path / 'missing.eam'))
