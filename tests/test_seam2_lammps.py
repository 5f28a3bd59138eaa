"""Seam 2, LAMMPS flavour: libatomistica_lammps.so exports what the reference's
src/lammps/pair_style/pair_atomistica.cpp links against; tests/lammps/pair_harness.cpp restates that
pair style's call sequence without LAMMPS (class lookup, particles_set_element, set_pointers with
the address-0 neighbour array and intptr_t seed/last, energy_and_forces into the live force array,
virial[6] -= ..., Voigt-6 per-atom virial).  A periodic system is unfolded into owned + ghost atoms
the way LAMMPS hands it over (REQ_FULL | REQ_GHOST) and must reproduce the periodic oracle.

Every case runs twice: through that restatement ('harness') and through the reference's UNMODIFIED
pair_atomistica.cpp ('pairstyle'), compiled from the reference tree against the stand-in LAMMPS headers of
tests/lammps/stub/ and driven like LAMMPS drives a pair style (settings, coeff, init_style, init_one,
compute; tests/lammps/pair_driver.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from atomistica_b200 import io as aio, structures as S
from atomistica_b200.seam2 import build as s2build
from conftest import load_npz


@pytest.fixture(scope='module')
def libs():
    if os.path.isdir('/root/reference/src/lammps'):
        s2build.build_lammps()
    lib, har = s2build.lammps_paths()
    if not (os.path.exists(lib) and os.path.exists(har)):
        pytest.skip('LAMMPS-flavour shim not built (needs /root/reference at build time)')
    return lib, har


def test_exports_the_symbols_of_the_pair_style(libs):
    out = subprocess.run(['nm', '-D', '--defined-only', libs[0]], capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in out.splitlines() if l.strip()}
    for sym in ('potential_classes', 'particles_new', 'particles_free', 'particles_init', 'particles_del',
                'particles_set_element', 'particles_set_pointers', 'particles_get_interaction_range',
                'particles_get_border', 'neighbors_new', 'neighbors_free', 'neighbors_init', 'neighbors_del',
                'neighbors_set_pointers', 'neighbors_get_cutoff', 'neighbors_dump_cutoffs', 'get_full_error_string',
                'atomistica_startup', 'atomistica_shutdown', 'ptrdict_read', 'ptrdict_cleanup'):
        assert sym in have, sym


def test_reference_pair_style_links_against_the_shim(libs):
    """the reference's unmodified pair_atomistica.cpp, compiled against stand-in LAMMPS headers, finds every
    symbol it needs in libatomistica_lammps.so (linked with --no-undefined) and resolves at load time"""
    ps = s2build.pairstyle_path()
    if not os.path.exists(ps):
        pytest.skip('the reference pair style was not built (needs /root/reference at build time)')
    out = subprocess.run(['nm', '-D', ps], capture_output=True, text=True).stdout
    need = {l.split()[-1] for l in out.splitlines() if ' U ' in l}
    have = {l.split()[-1] for l in subprocess.run(['nm', '-D', '--defined-only', libs[0]], capture_output=True,
                                                  text=True).stdout.splitlines() if l.strip()}
    ours = {s for s in need if s.startswith(('particles_', 'neighbors_', 'potential_classes', 'ptrdict_',
                                             'atomistica_', 'get_full_error_string'))}
    assert {'particles_set_pointers', 'neighbors_set_pointers', 'potential_classes', 'ptrdict_read',
            'atomistica_startup'} <= ours
    assert ours <= have, ours - have
    defined = subprocess.run(['nm', '-D', '--defined-only', ps], capture_output=True, text=True).stdout
    assert 'PairAtomistica' in defined and 'get_atomistica_pair_style_git_ident' in defined
    C.CDLL(ps)


def test_pair_style_owns_the_ptrdict_section(libs):
    """~PairAtomistica (pair_atomistica.cpp:137-142) cleans the ptrdict section itself and then calls del and
    free_instance: the LAMMPS-flavour instances must not clean it again (found by running the reference's own
    pair style: glibc aborted on the double free).  Runs in a child process, without a GPU."""
    ps = s2build.pairstyle_path()
    if not os.path.exists(ps):
        pytest.skip('the reference pair style was not built (needs /root/reference at build time)')
    code = ("import ctypes as C\n"
            "lib = C.CDLL(%r)\n"
            "for name in ('Tersoff', 'TersoffScr', 'Kumagai', 'KumagaiScr', 'Brenner', 'BrennerScr', 'TabulatedAlloyEAM',\n"
            "             'Rebo2', 'Juslin', 'JuslinScr'):\n"
            "    assert lib.lmp_pair_ownership(name.encode()) == 0, name\n"
            "print('ok')\n" % ps)
    r = subprocess.run([os.sys.executable, '-c', code], capture_output=True, text=True, env=dict(os.environ, MALLOC_CHECK_='3'))
    assert r.returncode == 0 and 'ok' in r.stdout, (r.returncode, r.stderr[-500:])


def _unfold(atoms, cutoff, shell):
    L = np.diag(atoms.cell)
    pos, sym, img = [atoms.positions], [list(atoms.symbols)], [np.arange(len(atoms))]
    w = shell * cutoff
    rng = [range(-int(np.ceil(w / L[k])), int(np.ceil(w / L[k])) + 1) for k in range(3)]
    for i in rng[0]:
        for j in rng[1]:
            for k in rng[2]:
                if (i, j, k) == (0, 0, 0):
                    continue
                p = atoms.positions + np.array([i, j, k]) * L
                m = np.all((p > -w) & (p < L + w), axis=1)
                pos.append(p[m]); sym.append([s for s, t in zip(atoms.symbols, m) if t]); img.append(np.nonzero(m)[0])
    pos = np.concatenate(pos); sym = sum(sym, []); img = np.concatenate(img)
    lists = []
    for i in range(len(pos)):
        d2 = ((pos - pos[i]) ** 2).sum(axis=1)
        nb = np.nonzero(d2 < cutoff * cutoff)[0]
        lists.append(np.ascontiguousarray(nb[nb != i], dtype=np.int32))
    return pos, sym, img, lists


DRIVERS = ('harness', 'pairstyle')


def _run(libs, name, atoms, cutoff, shell, types, param_file=None, eatom=True, vatom=True, ncalls=1, f0=None,
         driver='harness'):
    if driver == 'pairstyle':
        if not os.path.exists(s2build.pairstyle_path()):
            pytest.skip('the reference pair style was not built (needs /root/reference at build time)')
        har = C.CDLL(s2build.pairstyle_path())
        entry = har.lmp_pair_run
    else:
        har = C.CDLL(libs[1])
        entry = har.lmp_harness_run
    pos, sym, img, lists = _unfold(atoms, cutoff, shell)
    nall, nlocal = len(pos), len(atoms)
    x = np.ascontiguousarray(pos)
    tag = np.ascontiguousarray(img + 1, dtype=np.int32)
    typ = np.array([types.index(s) + 1 for s in sym], dtype=np.int32)
    ilist = np.arange(nall, dtype=np.int32)
    numneigh = np.array([len(l) for l in lists], dtype=np.int32)
    first = (C.POINTER(C.c_int) * nall)(*[l.ctypes.data_as(C.POINTER(C.c_int)) for l in lists])
    f = np.zeros((nall, 3)) if f0 is None else np.ascontiguousarray(f0(nall))
    ea, va, out = np.zeros(nall), np.zeros((nall, 6)), np.zeros(16)
    err = C.create_string_buffer(10000)
    syms = (C.c_char_p * len(types))(*[t.encode() for t in types])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = entry(name.encode(), (param_file or '').encode(), len(types), syms, nall, nlocal, ip(tag), ip(typ),
               dp(x), nall, ip(ilist), ip(numneigh), first, int(eatom), int(vatom), ncalls, dp(f), dp(ea),
               dp(va), dp(out), err)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors='replace'))
    return dict(eng=out[0], virial=out[1:7].copy(), rcghost=out[7], rc=out[8], rc_last=out[9], rc_cross=out[10], f=f, eatom=ea, vatom=va, nlocal=nlocal,
                img=img)


def test_no_cpu_fallback(libs):
    if os.path.exists('/dev/nvidia0'):
        pytest.skip('a GPU is present')
    a = S.diamond('Si', 5.432, (1, 1, 1))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_seam2_lammps as t, atomistica_b200.structures as S\n"
            "a = S.diamond('Si', 5.432, (1, 1, 1))\n"
            "t._run(%r, 'Tersoff', a, 3.0, 2, ['Si'])\n" % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)),
                                                          tuple(libs)))
    r = subprocess.run([os.sys.executable, '-c', code], capture_output=True, text=True)
    assert r.returncode != 0            # lammps_<pot>_init has no error argument: the reference aborts as well
    assert 'No CUDA device' in (r.stderr + r.stdout) or r.returncode < 0


def _voigt_minus(w):
    """virial[0..5] -= wpot(1,1), (2,2), (3,3), sym(2,1), sym(3,1), sym(3,2) (pair_atomistica.cpp:554-559)"""
    return -np.array([w[0, 0], w[1, 1], w[2, 2], 0.5 * (w[1, 0] + w[0, 1]), 0.5 * (w[2, 0] + w[0, 2]), 0.5 * (w[2, 1] + w[1, 2])])


@pytest.mark.gpu
@pytest.mark.parametrize('driver', DRIVERS)
def test_tersoff_through_the_pair_style_sequence(libs, driver):
    import oracle
    from atomistica_b200 import parameters as P
    a = S.diamond('Si', 5.432, (3, 3, 3))
    a.rattle(0.08, seed=4)
    db = P.complete('Tersoff', None)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.TERSOFF, db), a.positions, a.cell, onl, el, per_at=True)
    # the live force array of LAMMPS is not zero on entry: the potential ADDS (tls_reduce, bop_kernel.f90:1615-1620)
    r = _run(libs, 'Tersoff', a, 3.0 + 0.3, 2, ['Si'], f0=lambda n: np.full((n, 3), 0.25), ncalls=2, driver=driver)
    n = r['nlocal']
    assert abs(r['rc'] - 3.0) < 1e-12 and abs(r['rcghost'] - 6.0) < 1e-12      # list cutoff, 2 x cutoff ghost shell
    assert abs(r['eng'] - 2 * o['epot']) <= 1e-10 * abs(2 * o['epot'])          # two calls accumulate eng_vdwl
    fs = max(1.0, np.abs(o['f']).max())
    assert np.abs(r['f'][:n] - 0.25 - 2 * o['f']).max() <= 1e-10 * fs            # added twice onto the start values
    assert np.all(r['f'][n:] == 0.25)                                             # ghost rows untouched (gather)
    assert np.abs(r['virial'] - 2 * _voigt_minus(o['wpot'])).max() <= 1e-10 * max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
    assert np.abs(r['eatom'][:n] - 2 * o['epot_per_at']).max() <= 1e-10 * np.abs(o['epot_per_at']).max()
    # Voigt-6 per-atom virial with the minus sign of the LAMMPS build (macros.inc:202): sums to the total
    w = o['wpot_per_at'].sum(axis=0)
    tot = -np.array([w[0, 0], w[1, 1], w[2, 2], w[1, 0], w[2, 0], w[2, 1]])
    assert np.abs(r['vatom'][:n].sum(axis=0) - 2 * tot).max() <= 1e-9 * max(1.0, np.abs(w).max(), abs(o['epot']))
    wpa = o['wpot_per_at']
    ref6 = -np.stack([wpa[:, 0, 0], wpa[:, 1, 1], wpa[:, 2, 2], wpa[:, 1, 0], wpa[:, 2, 0], wpa[:, 2, 1]], axis=1)
    assert np.abs(r['vatom'][:n] - 2 * ref6).max() <= 1e-9 * max(1.0, np.abs(ref6).max())


@pytest.mark.gpu
@pytest.mark.parametrize('driver', DRIVERS)
def test_two_types_eam_and_rebo2(libs, tmp_path, driver):
    import oracle
    from atomistica_b200 import parameters as P
    # Brenner SiC: two LAMMPS types mapped to elements by particles_set_element
    a = S.b3(['Si', 'C'], 4.3596, (3, 3, 3))
    a.rattle(0.08, seed=5)
    db = P.complete('Brenner', None)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.BRENNER, db), a.positions, a.cell, onl, el)
    r = _run(libs, 'Brenner', a, max(db['r2']) + 0.3, 2, ['Si', 'C'], vatom=False, driver=driver)
    # neighbors_get_cutoff(i, j): the cutoff of THAT pair of types (lammps_neighbors.f90:223-251), el = ['C', 'Si']
    assert abs(r['rc'] - db['r2'][2]) < 1e-12 and abs(r['rc_last'] - db['r2'][0]) < 1e-12   # Si-Si, C-C
    assert abs(r['rc_cross'] - db['r2'][1]) < 1e-12                                          # Si-C
    assert abs(r['eng'] - o['epot']) <= 1e-10 * abs(o['epot'])
    assert np.abs(r['f'][:r['nlocal']] - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())
    # TabulatedAlloyEAM with its setfl file given in a ptrdict parameter file (pair_style atomistica <name> <file>)
    setfl = load_npz('cu_mishin1_setfl.npz')
    fn = str(tmp_path / 'Cu.eam.alloy')
    aio.write_setfl(fn, setfl)
    par = tmp_path / 'eam.dat'
    par.write_text('TabulatedAlloyEAM {\n  fn = "%s";\n};\n' % fn)
    cu = S.fcc('Cu', 3.615, (4, 4, 4))
    cu.rattle(0.05, seed=3)
    eam = oracle.EAM(setfl)
    onl = oracle.neighbor_list(cu.positions, cu.cell, cu.pbc, eam.cutoff, 200)
    o = eam.energy_and_forces(cu.positions, cu.cell, onl, eam.eldb(cu.symbols))
    r = _run(libs, 'TabulatedAlloyEAM', cu, eam.cutoff + 0.3, 2, ['Cu'], param_file=str(par), vatom=False, driver=driver)
    assert abs(r['eng'] - o['epot']) <= 1e-10 * abs(o['epot'])
    assert np.abs(r['f'][:r['nlocal']] - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())
    # REBO2: every bond once (the reference decides by atom tag, bop_kernel_rebo2.f90:1329), 5-bond ghost shell
    c = S.diamond('C', 3.6, (3, 3, 3))
    c.rattle(0.05, seed=8)
    rb = oracle.Rebo2()
    onl = oracle.neighbor_list(c.positions, c.cell, c.pbc, 2.0, 50)
    o = rb.energy_and_forces(c.positions, c.cell, onl, rb.ktyp(c.symbols))
    r = _run(libs, 'Rebo2', c, 2.0, 5, ['C'], vatom=False, driver=driver)
    assert abs(r['rcghost'] - 10.0) < 1e-12
    assert abs(r['eng'] - o['epot']) <= 1e-10 * abs(o['epot'])
    assert np.abs(r['f'][:r['nlocal']] - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())


@pytest.mark.gpu
@pytest.mark.parametrize('driver', DRIVERS)
def test_juslin_through_the_pair_style_sequence(libs, driver):
    """Juslin W-C (non-symmetric pair index, rows mirrored by init) on an unfolded B1 crystal"""
    import oracle
    from atomistica_b200 import parameters as P
    a = S.b1(['W', 'C'], 4.38, (3, 3, 3))
    a.rattle(0.08, seed=6)
    db = P.complete_juslin(None)
    nel = len(db['el'])
    idx = [db['el'].index(s) for s in ('W', 'C')]
    cutoff = max(db['r2'][j + i * nel] for i in idx for j in idx)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 200)
    el = np.array([db['el'].index(s) + 1 for s in a.symbols], dtype=np.int32)
    o = oracle.bop_energy_and_forces(oracle.bop_params(oracle.JUSLIN, db), a.positions, a.cell, onl, el)
    r = _run(libs, 'Juslin', a, cutoff + 0.3, 2, ['W', 'C'], vatom=False, driver=driver)
    assert abs(r['eng'] - o['epot']) <= 1e-10 * abs(o['epot'])
    assert np.abs(r['f'][:r['nlocal']] - o['f']).max() <= 1e-10 * max(1.0, np.abs(o['f']).max())
    assert np.abs(r['virial'] - _voigt_minus(o['wpot'])).max() <= 1e-10 * max(1.0, np.abs(o['wpot']).max(), abs(o['epot']))
