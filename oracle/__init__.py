"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement of the Atomistica hot path on the CPU (plain C kernels in
oracle_*.c, init-time table/spline construction in numpy here).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package `atomistica_b200` never does.

Parity status: the Fortran reference cannot be built in the build container
(no Fortran compiler), so there is no oracle/_ref.  The oracle is pinned (1) by the
reference's own known-answer tests (tests/test_oracle_kat.py against
tests/golden/kat.json) and (2) by the reference's own Fortran kernels EXECUTED next to
it through a Fortran-subset translator (tests/test_func_vs_reference.py: neighbour lists
equal entry by entry; energies, forces, virials, per-atom and per-bond outputs of the
EAM, Tersoff / Kumagai / Brenner (plain and screened), Juslin and REBO2 / Rebo2Scr
kernels at <= 1e-12) -- not by output of a compiled reference binary.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRCS = ['oracle_neighbors.c', 'oracle_eam.c', 'oracle_bop.c', 'oracle_rebo2.c', 'oracle_pair.c']
_LIB = os.path.join(_HERE, 'liboracle.so')
_lib = None

PI = 3.14159265358979323846264338327950288


def build(force=False):
    """gcc -O2 -ffp-contract=off: every multiply/add rounded separately, like the
    reference's gfortran x86-64 baseline build (no FMA contraction)."""
    srcs = [os.path.join(_HERE, s) for s in _SRCS]
    hdr = os.path.join(_HERE, 'atomistica_oracle.h')
    if not force and os.path.exists(_LIB):
        newest = max(os.path.getmtime(p) for p in srcs + [hdr] if os.path.exists(p))
        if not all(os.path.exists(p) for p in srcs) or os.path.getmtime(_LIB) >= newest:
            return _LIB
    cmd = ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-std=c99',
           '-o', _LIB] + srcs + ['-lm']
    subprocess.check_call(cmd)
    return _LIB


_strict = None
_fast = None


def lib():
    global _lib, _strict
    if _lib is None:
        build()
        _strict = C.CDLL(_LIB)
        _strict.orc_nl_build.restype = C.c_long
        _lib = _strict
    return _lib


def use_fast(on=True):
    """bench.py's CPU legs only: switch to a copy of the oracle compiled ON THIS MACHINE with
    -O3 -march=native (what SURVEY 8(d) asks the timed CPU path to use).  The known-answer tests
    always run the strict -O2 -ffp-contract=off build.  Falls back to the strict build when the
    compilation is not possible.  Returns the flags in use."""
    global _lib, _fast
    lib()
    if not on:
        _lib = _strict
        return 'gcc -O2 -fopenmp -ffp-contract=off'
    if _fast is None:
        import tempfile
        out = os.path.join(tempfile.gettempdir(), 'atx_oracle_fast_%d' % os.getuid())
        try:
            os.makedirs(out, exist_ok=True)
            so = os.path.join(out, 'liboracle_fast.so')
            srcs = [os.path.join(_HERE, s) for s in _SRCS]
            subprocess.check_call(['gcc', '-O3', '-march=native', '-fopenmp', '-fPIC', '-shared', '-std=c99',
                                   '-I' + _HERE, '-o', so] + srcs + ['-lm'], stderr=subprocess.DEVNULL)
            _fast = C.CDLL(so)
            _fast.orc_nl_build.restype = C.c_long
        except Exception:
            _fast = False
    if _fast:
        _lib = _fast
        return 'gcc -O3 -march=native -fopenmp'
    return 'gcc -O2 -fopenmp -ffp-contract=off'


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


# ----------------------------------------------------------------------------
# particles helpers (python_particles.f90:286-346)
# ----------------------------------------------------------------------------

def abox_from_cell(cell):
    """ASE cell rows are the cell vectors; the C-contiguous buffer read column-major
    gives Abox(:,k) = cell vector k (src/python/c/particles.c:436-461)."""
    cell = np.asarray(cell, dtype=np.float64)
    if cell.shape == (3,):
        cell = np.diag(cell)
    return np.ascontiguousarray(cell).ravel().copy()


def bbox_from_abox(abox):
    """Bbox = Abox^-1 via LAPACK dgesv (f_linearalgebra.f90:599-637); numpy.linalg.solve is dgesv."""
    A = abox.reshape(3, 3).T  # A[i,j] = Abox(i,j)
    B = np.linalg.solve(A, np.eye(3))
    return np.ascontiguousarray(B.T).ravel().copy()  # back to column-major


# ----------------------------------------------------------------------------
# neighbour list
# ----------------------------------------------------------------------------

class NeighborList:
    pass


def neighbor_list(r, cell, pbc, cutoff, avgn=100):
    r = np.ascontiguousarray(r, dtype=np.float64)
    nat = len(r)
    abox = abox_from_cell(cell)
    bbox = bbox_from_abox(abox)
    pbc = np.ascontiguousarray(np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).astype(np.int32))
    cap = max(nat * avgn, 1)
    nl = NeighborList()
    nl.seed = np.zeros(nat + 1, dtype=np.intp)
    nl.last = np.zeros(nat + 1, dtype=np.intp)
    nl.neighbors = np.zeros(cap, dtype=np.int32)
    nl.dc = np.zeros((cap, 3), dtype=np.int32)
    n = lib().orc_nl_build(C.c_int(nat), _p(r), _p(abox), _p(bbox), _p(pbc, C.c_int), C.c_double(cutoff),
                           C.c_long(cap), _p(nl.seed, C.c_ssize_t), _p(nl.last, C.c_ssize_t),
                           _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int))
    if n < 0:
        raise RuntimeError('Neighbor list overflow.')
    nl.npairs = int(n)
    nl.abox = abox
    nl.cutoff = cutoff
    return nl


def pairs(nl, nat):
    """(i, j, dc) rows (0-based i, j) in list order."""
    i = np.repeat(np.arange(nat), (nl.last[:nat] - nl.seed[:nat] + 1))
    slots = np.concatenate([np.arange(nl.seed[k] - 1, nl.last[k]) for k in range(nat)]) if nat else np.zeros(0, int)
    return i, nl.neighbors[slots] - 1, nl.dc[slots], slots


# ----------------------------------------------------------------------------
# simple_spline (simple_spline.f90:127-195)
# ----------------------------------------------------------------------------

class Spline(C.Structure):
    _fields_ = [('n', C.c_int), ('x0', C.c_double), ('dx', C.c_double), ('cut', C.c_double)] + \
        [(k, C.POINTER(C.c_double)) for k in ('y', 'c1', 'c2', 'c3', 'd1', 'd2', 'd3')]


def spline_init(n, x0, dx, y):
    y = np.array(y, dtype=np.float64)
    assert len(y) == n
    sig = 0.5
    d2y = np.zeros(n)
    u = np.zeros(n)
    yl = y.tolist()
    d2 = [0.0] * n
    ul = [0.0] * n
    for i in range(1, n - 1):
        p = sig * d2[i - 1] + 2
        d2[i] = (sig - 1) / p
        ul[i] = (6.0 * ((yl[i + 1] - yl[i]) / dx - (yl[i] - yl[i - 1]) / dx) / (2 * dx) - sig * ul[i - 1]) / p
    qn = 0.0
    un = 0.0
    d2[n - 1] = (un - qn * ul[n - 2]) / (qn * d2[n - 2] + 1.0)
    for k in range(n - 2, -1, -1):
        d2[k] = d2[k] * d2[k + 1] + ul[k]
    d2y = np.array(d2)
    c1 = y[1:] - y[:-1] - (2 * d2y[:-1] + d2y[1:]) * dx ** 2 / 6
    c2 = d2y[:-1] * dx ** 2 / 2
    c3 = (d2y[1:] - d2y[:-1]) * dx ** 2 / 6
    s = dict(n=n, x0=x0, dx=dx, cut=x0 + dx * (n - 1), y=y, d2y=d2y, c1=c1, c2=c2, c3=c3,
             d1=c1 / dx, d2=2 * c2 / dx, d3=3 * c3 / dx)
    return s


def spline_scale_y(s, fac):
    for k in ('y', 'd2y', 'c1', 'c2', 'c3', 'd1', 'd2', 'd3'):
        s[k] = fac * s[k]
    return s


def _spline_struct(s):
    st = Spline()
    st.n, st.x0, st.dx, st.cut = s['n'], s['x0'], s['dx'], s['cut']
    for k in ('y', 'c1', 'c2', 'c3', 'd1', 'd2', 'd3'):
        s[k] = np.ascontiguousarray(s[k], dtype=np.float64)
        setattr(st, k, _p(s[k]))
    return st


def spline_eval(s, x, extrapolate=False):
    """f_and_df (simple_spline.f90:536-614), numpy scalar version for unit tests."""
    n, x0, dx = s['n'], s['x0'], s['dx']
    if extrapolate:
        xf = (x - x0) / dx + 1
        i = int(np.floor(xf))
        i = min(max(i, 1), n - 1)
    else:
        if x == s['cut']:
            xf, i = float(n), n - 1
        else:
            xf = (x - x0) / dx + 1
            i = int(np.floor(xf))
        if i < 1 or i >= n:
            raise ValueError('x outside of the defined interval')
    B = xf - i
    i -= 1
    f = s['y'][i] + B * (s['c1'][i] + B * (s['c2'][i] + B * s['c3'][i]))
    df = s['d1'][i] + B * (s['d2'][i] + B * s['d3'][i])
    return f, df


# ----------------------------------------------------------------------------
# tabulated alloy EAM (tabulated_alloy_eam.f90:147-259 init; kernel in C)
# ----------------------------------------------------------------------------

def set_threads(n):
    """OpenMP threads of the EAM kernel and of the list build (bench.py's CPU legs); 1 = the serial
    code the KATs pin"""
    lib().orc_eam_set_threads(int(n))
    lib().orc_nl_set_threads(int(n))


class EAM:
    def __init__(self, setfl):
        nel = len(setfl['names'])
        self.names = [str(x) for x in setfl['names']]
        self.cutoff = float(setfl['cutoff'])
        nF, dF, nr, dr = int(setfl['nF']), float(setfl['dF']), int(setfl['nr']), float(setfl['dr'])
        self.fF = [spline_init(nF, 0.0, dF, setfl['F'][i]) for i in range(nel)]
        pad = np.zeros(2)
        self.frho = [spline_init(nr + 2, 0.0, dr, np.concatenate([setfl['rho'][i], pad])) for i in range(nel)]
        self.fphi = [[None] * nel for _ in range(nel)]
        k = 0
        for i in range(nel):
            for j in range(i + 1):
                s = spline_scale_y(spline_init(nr + 2, 0.0, dr, np.concatenate([setfl['rphi'][k], pad])), 0.5)
                self.fphi[i][j] = s
                self.fphi[j][i] = s
                k += 1
        self.nel = nel

    def eldb(self, symbols):
        return np.array([self.names.index(s) + 1 if s in self.names else -1 for s in symbols], dtype=np.int32)

    def energy_and_forces(self, r, cell, nl, eldb, mask=None, per_at=False):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nat = len(r)
        abox = abox_from_cell(cell)
        ndb = self.nel
        fF = (Spline * ndb)(*[_spline_struct(s) for s in self.fF])
        frho = (Spline * ndb)(*[_spline_struct(s) for s in self.frho])
        fphi = (Spline * (ndb * ndb))(*[_spline_struct(self.fphi[i][j]) for j in range(ndb) for i in range(ndb)])
        epot = C.c_double(0.0)
        f = np.zeros((nat, 3))
        wpot = np.zeros(9)
        epa = np.zeros(nat) if per_at else None
        wpa = np.zeros((nat, 9)) if per_at else None
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
        eldb = np.ascontiguousarray(eldb, dtype=np.int32)
        err = lib().orc_eam_energy_and_forces(
            C.c_int(nat), C.c_int(nat), _p(r), _p(abox), _p(eldb, C.c_int), _p(nl.seed, C.c_ssize_t),
            _p(nl.last, C.c_ssize_t), _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int), C.c_int(ndb), fF, frho, fphi,
            C.c_double(self.cutoff), _p(m, C.c_int), C.byref(epot), _p(f), _p(wpot), _p(epa), _p(wpa))
        if err:
            raise RuntimeError('spline argument outside of the defined interval')
        out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
        if per_at:
            out['epot_per_at'] = epa
            out['wpot_per_at'] = wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy()
        return out


HARTREE, BOHR = 27.2113961, 0.529177249       # src/support/Units.f90:74-76


class EAMFuncfl:
    """TabulatedEAM: single-element funcfl tables (tabulated_eam.f90:141-214)"""

    def __init__(self, tab, elements='*'):
        self.name = str(tab['name'])
        self.elements = elements
        self.cutoff = float(tab['cutoff'])
        nF, dF, nr, dr = int(tab['nF']), float(tab['dF']), int(tab['nr']), float(tab['dr'])
        self.fF = spline_init(nF, 0.0, dF, tab['F'])
        self.fZ = spline_scale_y(spline_init(nr, 0.0, dr, tab['Z']), np.sqrt(0.5 * HARTREE * BOHR))
        self.frho = spline_init(nr, 0.0, dr, tab['rho'])

    def energy_and_forces(self, r, cell, nl, symbols, per_at=False):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nat = len(r)
        abox = abox_from_cell(cell)
        sel = [x.strip() for x in self.elements.split(',')]
        inn = np.array([1 if (self.elements.strip() == '*' or s in sel) else 0 for s in symbols], dtype=np.int32)
        epot = C.c_double(0.0)
        f = np.zeros((nat, 3))
        wpot = np.zeros(9)
        epa = np.zeros(nat) if per_at else None
        sF, sZ, sR = _spline_struct(self.fF), _spline_struct(self.fZ), _spline_struct(self.frho)
        err = lib().orc_eam_funcfl_energy_and_forces(
            C.c_int(nat), _p(r), _p(abox), _p(inn, C.c_int), _p(nl.seed, C.c_ssize_t), _p(nl.last, C.c_ssize_t),
            _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int), C.byref(sF), C.byref(sZ), C.byref(sR),
            C.c_double(self.cutoff), C.byref(epot), _p(f), _p(wpot), _p(epa))
        assert err == 0
        out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
        if per_at:
            out['epot_per_at'] = epa
        return out


# ----------------------------------------------------------------------------
# generic bond-order potentials
# ----------------------------------------------------------------------------

TERSOFF, KUMAGAI, BRENNER, JUSLIN = 1, 2, 3, 4


class BopParams(C.Structure):
    _fields_ = [('kind', C.c_int), ('nel', C.c_int), ('pp', (C.c_double * 9) * 12), ('ep', (C.c_double * 3) * 8),
                ('ip', C.c_int * 9), ('r1', C.c_double * 9), ('r2', C.c_double * 9),
                ('t_alpha', C.c_double * 27), ('t_omega', C.c_double * 27), ('t_m', C.c_int * 27)]


_PAIR_ROWS = {
    TERSOFF: ['A', 'B', 'xi', 'lambda', 'mu', 'omega', 'mubo'],
    KUMAGAI: ['A', 'B', 'lambda1', 'lambda2', 'alpha'],
    BRENNER: ['D0', 'r0', 'S', 'beta', 'gamma', 'c', 'd', 'h', 'mu', 'n'],
    JUSLIN: ['D0', 'r0', 'S', 'beta', 'gamma', 'c', 'd', 'h', None, 'n'],
}
_EL_ROWS = {
    TERSOFF: ['beta', 'n', 'c', 'd', 'h'],
    KUMAGAI: ['eta', 'delta', 'c1', 'c2', 'c3', 'c4', 'c5', 'h'],
    BRENNER: [],
    JUSLIN: [],
}
_INT_ROW = {TERSOFF: 'm', KUMAGAI: 'beta', BRENNER: 'm', JUSLIN: None}


def bop_params(kind, db):
    """db: dict of lists in the layout of atomistica/parameters.py (pair lists in PAIR_INDEX order;
    Juslin: PAIR_INDEX_NS / TRIPLET_INDEX_NS order, already mirrored like juslin_module.f90:283-312)."""
    p = BopParams()
    p.kind = kind
    nel = len(db['el'])
    p.nel = nel
    npairs = nel * nel if kind == JUSLIN else nel * (nel + 1) // 2
    for row, key in enumerate(_PAIR_ROWS[kind]):
        for k in range(npairs):
            p.pp[row][k] = float(db[key][k]) if key is not None else 0.0
    for row, key in enumerate(_EL_ROWS[kind]):
        for k in range(nel):
            p.ep[row][k] = float(db[key][k])
    for k in range(npairs):
        p.ip[k] = int(db[_INT_ROW[kind]][k]) if _INT_ROW[kind] is not None else 1
        p.r1[k] = float(db['r1'][k])
        p.r2[k] = float(db['r2'][k])
    if kind == JUSLIN:
        for k in range(nel ** 3):
            p.t_alpha[k] = float(db['alpha'][k])
            p.t_omega[k] = float(db['omega'][k])
            p.t_m[k] = int(db['m'][k])
    return p


def _per_bond_arrays(nl, per_bond):
    if not per_bond:
        return None, None, None
    n = len(nl.neighbors)
    return np.zeros(n), np.zeros((n, 3)), np.zeros((n, 9))


PAIR_LJCUT, PAIR_HARMONIC, PAIR_DOUBLE_HARMONIC, PAIR_BORN_MAYER, PAIR_R6 = 1, 2, 3, 4, 5


def element_ids(symbols):
    """compact 1-based particle element ids, here in order of first appearance (the reference numbers them by
    ascending atomic number, python_particles.f90:617-658; the ids are internal -- filters and el2db maps are built
    from the same numbering, so any consistent one gives the same energies); returns (ids per atom, symbols by id-1)"""
    order = []
    for s in symbols:
        if s not in order:
            order.append(s)
    return np.array([order.index(s) + 1 for s in symbols], dtype=np.int32), order


def element_filter(spec, order):
    """filter_from_string (src/core/filter.f90:55-120): '*' or comma-separated symbols -> bit mask"""
    if spec.strip() == '*':
        return sum(1 << (k + 1) for k in range(len(order)))
    return sum(1 << (order.index(s.strip()) + 1) for s in spec.split(',') if s.strip() in order)


def pair_energy_and_forces(kind, par, r, cell, nl, symbols, el1='*', el2='*', shift=False, mask=None, per_at=False):
    r = np.ascontiguousarray(r, dtype=np.float64)
    nat = len(r)
    abox = abox_from_cell(cell)
    el, order = element_ids(symbols)
    par = np.ascontiguousarray(par, dtype=np.float64)
    epot = C.c_double(0.0)
    f = np.zeros((nat, 3))
    wpot = np.zeros(9)
    epa = np.zeros(nat) if per_at else None
    wpa = np.zeros((nat, 9)) if per_at else None
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
    err = lib().orc_pair_energy_and_forces(
        C.c_int(kind), _p(par), C.c_int(1 if shift else 0), C.c_int(nat), _p(r), _p(abox), _p(el, C.c_int),
        C.c_int(element_filter(el1, order)), C.c_int(element_filter(el2, order)), _p(nl.seed, C.c_ssize_t),
        _p(nl.last, C.c_ssize_t), _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int), _p(m, C.c_int), C.byref(epot),
        _p(f), _p(wpot), _p(epa), _p(wpa))
    assert err == 0, err
    out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
    if per_at:
        out['epot_per_at'] = epa
        out['wpot_per_at'] = wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy()
    return out


class BopScr(C.Structure):
    _fields_ = [(k, C.c_double * 9) for k in ('or1', 'or2', 'bor1', 'bor2', 'Cmin', 'Cmax')]


def bop_scr_params(db):
    """outer / bond-order cutoffs and screening bounds of a *__Scr parameter set"""
    s = BopScr()
    npairs = len(db['or1'])      # el (el + 1) / 2, or el x el for the Juslin sets
    for key in ('or1', 'or2', 'bor1', 'bor2', 'Cmin', 'Cmax'):
        for k in range(npairs):
            getattr(s, key)[k] = float(db[key][k])
    return s


def bop_energy_and_forces(params, r, cell, nl, el, mask=None, per_at=False, per_bond=False, scr=None):
    r = np.ascontiguousarray(r, dtype=np.float64)
    nat = len(r)
    abox = abox_from_cell(cell)
    el = np.ascontiguousarray(el, dtype=np.int32)
    epot = C.c_double(0.0)
    f = np.zeros((nat, 3))
    wpot = np.zeros(9)
    epa = np.zeros(nat) if per_at else None
    wpa = np.zeros((nat, 9)) if per_at else None
    epb, fpb, wpb = _per_bond_arrays(nl, per_bond)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
    tail = (C.c_int(nat), C.c_int(nat), _p(r), _p(abox), _p(el, C.c_int), _p(nl.seed, C.c_ssize_t),
            _p(nl.last, C.c_ssize_t), _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int), _p(m, C.c_int), C.byref(epot),
            _p(f), _p(wpot), _p(epa), _p(epb), _p(fpb), _p(wpa), _p(wpb))
    if scr is None:
        err = lib().orc_bop_energy_and_forces(C.byref(params), *tail)
    else:
        err = lib().orc_bop_scr_energy_and_forces(C.byref(params), C.byref(scr), *tail)
    assert err == 0
    out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
    if per_at:
        out['epot_per_at'] = epa
        out['wpot_per_at'] = wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy()
    if per_bond:
        out['epot_per_bond'] = epb
        out['f_per_bond'] = fpb
        out['wpot_per_bond'] = wpb.reshape(-1, 3, 3).transpose(0, 2, 1).copy()
    return out


# ----------------------------------------------------------------------------
# REBO2: tables (table2d.f90:84-226, table3d.f90:85-284), g spline (rebo2_db.f90:405-524),
# constants (rebo2_type.f90:48-394, rebo2_db.f90:81-303), default tables
# (rebo2_default_tables.f90)
# ----------------------------------------------------------------------------

class Table2d(C.Structure):
    _fields_ = [('nx', C.c_int), ('ny', C.c_int), ('coeff', C.POINTER(C.c_double))]


class Table3d(C.Structure):
    _fields_ = [('nx', C.c_int), ('ny', C.c_int), ('nz', C.c_int), ('coeff', C.POINTER(C.c_double))]


class Rebo2Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        'cc_B1', 'cc_B2', 'cc_B3', 'cc_beta1', 'cc_beta2', 'cc_beta3', 'cc_Q', 'cc_A', 'cc_alpha',
        'ch_B1', 'ch_beta1', 'ch_Q', 'ch_A', 'ch_alpha', 'hh_B1', 'hh_beta1', 'hh_Q', 'hh_A', 'hh_alpha')] + [
        ('cc_g_theta', C.c_double * 6), ('cc_g1_coeff', C.c_double * 18), ('cc_g2_coeff', C.c_double * 18),
        ('spgh', C.c_double * 18), ('igh', C.c_int * 25), ('conalp', C.c_double), ('conear', C.c_double * 36),
        ('conpe', C.c_double * 3), ('conan', C.c_double * 3), ('conpf', C.c_double * 3),
        ('cut_in_l', C.c_double * 10), ('cut_in_h', C.c_double * 10), ('cut_in_h2', C.c_double * 10),
        ('with_dihedral', C.c_int), ('Fcc', Table3d), ('Fch', Table3d), ('Fhh', Table3d), ('Tcc', Table3d),
        ('Pcc', Table2d), ('Pch', Table2d)]


def _f32(x):
    """single-precision literal promoted to double"""
    return float(np.float32(x))


REBO2_DEFAULTS = dict(
    cc_B1=12388.79197798, cc_B2=17.56740646509, cc_B3=30.71493208065,
    cc_beta1=4.7204523127, cc_beta2=1.4332132499, cc_beta3=1.3826912506,
    cc_Q=0.3134602960833, cc_A=10953.544162170, cc_alpha=4.7465390606595,
    ch_B1=32.3551866587, ch_beta1=1.43445805925, ch_Q=0.340775728, ch_A=149.94098723, ch_alpha=4.10254983,
    hh_B1=29.632593, hh_beta1=1.71589217, hh_Q=0.370471487045, hh_A=32.817355747, hh_alpha=3.536298648,
    hhh_lambda=4.0, cc_re=1.4, ch_re=1.09, hh_re=0.7415886997,
    cc_in_r1=1.70, cc_in_r2=2.00, ch_r1=1.30, ch_r2=1.80, hh_r1=1.10, hh_r2=1.70,
    with_dihedral=False,
)
# rebo2_type.f90:104-108 -- g1/dg1/d2g1/g2 are default-real (single precision) array constructors
CC_G_THETA = [-1.0, -1.0 / 2, -1.0 / 3, 0.0, 1.0 / 2, 1.0]
CC_G_G1 = [_f32(x) for x in (-0.01, 0.05280, 0.09733, 0.37545, 2.0014, 8.0)]
CC_G_DG1 = [_f32(x) for x in (0.10400, 0.17000, 0.40000, 0.0, 0.0, 0.0)]
CC_G_D2G1 = [_f32(x) for x in (0.00000, 0.37000, 1.98000, 0.0, 0.0, 0.0)]
CC_G_G2 = [_f32(x) for x in (0.0, 0.0, 0.09733, 0.271856, 0.416335, 1.0)]
SPGH = [270.467795364007301, 1549.701314596994564, 3781.927258631323866, 4582.337619544424228,
        2721.538161662818368, 630.658598136730774, 16.956325544514659, -21.059084522755980,
        -102.394184748124742, -210.527926707779059, -229.759473570467513, -94.968528666251945,
        19.065031149937783, 2.017732531534021, -2.566444502991983, 3.291353893907436,
        -2.653536801884563, 0.837650930130006]
IGH = [3] * 18 + [2] * 4 + [1] * 3


def rebo2_default_tables():
    """rebo2_default_tables.f90:32-336; arrays indexed [i, j, k] like the Fortran (0:4,0:4,0:9)."""
    F = np.zeros((5, 5, 10)); dFdi = np.zeros((5, 5, 10)); dFdj = np.zeros((5, 5, 10)); dFdk = np.zeros((5, 5, 10))
    F[1, 1, 0] = 0.105000
    F[1, 1, 1] = -0.0041775
    F[1, 1, 2:9] = -0.0160856
    F[2, 2, 0] = 0.09444957
    F[2, 2, 1] = 0.02200000
    F[2, 2, 2] = 0.03970587
    F[2, 2, 3] = 0.03308822
    F[2, 2, 4] = 0.02647058
    F[2, 2, 5] = 0.01985293
    F[2, 2, 6] = 0.01323529
    F[2, 2, 7] = 0.00661764
    F[2, 2, 8] = 0.0
    F[0, 1, 0] = 0.04338699
    F[0, 1, 1] = 0.0099172158
    F[0, 1, 1:9] = 0.0099172158
    F[0, 2, 0] = 0.0493976637
    F[0, 2, 1] = -0.011942669
    F[0, 2, 2:9] = F[0, 1, 1]
    F[0, 3, 0:9] = -0.119798935
    F[0, 3, 0:2] = -0.119798935
    F[0, 3, 2:9] = F[0, 1, 1]
    F[1, 2, 0] = 0.0096495698
    F[1, 2, 1] = 0.030
    F[1, 2, 2] = -0.0200
    F[1, 2, 3] = -0.0233778774
    F[1, 2, 4] = -0.0267557548
    F[1, 2, 5:9] = -0.030133632
    F[1, 3, 1:9] = -0.124836752
    F[2, 3, 0:9] = -0.044709383
    for i in range(3, 8):
        F[2, 2, i] = F[2, 2, 2] + (i - 2) * (F[2, 2, 8] - F[2, 2, 2]) / 6
    for i in range(3, 5):
        F[1, 2, i] = F[1, 2, 2] + (i - 2) * (F[1, 2, 5] - F[1, 2, 2]) / 3
    dFdi[2, 1, 0] = -0.052500
    dFdi[2, 1, 4:9] = -0.054376
    dFdi[2, 3, 0] = 0.0
    dFdi[2, 3, 1:6] = 0.062418
    dFdk[2, 2, 3:8] = -0.006618
    dFdi[2, 3, 6:9] = 0.062418
    dFdk[1, 1, 1] = -0.060543
    dFdk[1, 2, 3] = -0.020044
    dFdk[1, 2, 4] = -0.020044
    for k in range(10):
        for i in range(4):
            for j in range(i + 1, 4):
                x = F[i, j, k] + F[j, i, k]; F[i, j, k] = x; F[j, i, k] = x
                x = dFdi[i, j, k] + dFdj[j, i, k]; dFdi[i, j, k] = x; dFdj[j, i, k] = x
                x = dFdi[j, i, k] + dFdj[i, j, k]; dFdi[j, i, k] = x; dFdj[i, j, k] = x
                x = dFdk[i, j, k] + dFdk[j, i, k]; dFdk[i, j, k] = x; dFdk[j, i, k] = x
    Fch = np.zeros((5, 5, 10))
    Fch[0, 2, 4:9] = -0.0090477875161288110
    Fch[1, 3, 0:9] = -0.213
    Fch[1, 2, 0:9] = -0.25
    Fch[1, 1, 0:9] = -0.5
    for k in range(10):
        for i in range(3):
            for j in range(i + 1, 4):
                x = Fch[i, j, k] + Fch[j, i, k]; Fch[i, j, k] = x; Fch[j, i, k] = x
    Fhh = np.zeros((5, 5, 10))
    Fhh[1, 1, 0] = 0.249831916
    Pcc = np.zeros((6, 6))
    Pcc[1, 1] = 0.003026697473481
    Pcc[2, 0] = 0.007860700254745
    Pcc[3, 0] = 0.016125364564267
    Pcc[1, 2] = 0.003179530830731
    Pcc[2, 1] = 0.006326248241119
    Pch = np.zeros((6, 6))
    Pch[1, 0] = 0.2093367328250380
    Pch[2, 0] = -0.064449615432525
    Pch[3, 0] = -0.303927546346162
    Pch[0, 1] = 0.01
    Pch[0, 2] = -0.1220421462782555
    Pch[1, 1] = -0.1251234006287090
    Pch[2, 1] = -0.298905245783
    Pch[0, 3] = -0.307584705066
    Pch[1, 2] = -0.3005291724067579
    Tcc = np.zeros((5, 5, 10))
    Tcc[2, 2, 0] = -0.070280085
    Tcc[2, 2, 1:9] = -0.00809675
    return dict(Fcc=F, dFdi=dFdi, dFdj=dFdj, dFdk=dFdk, Fch=Fch, Fhh=Fhh, Pcc=Pcc, Pch=Pch, Tcc=Tcc)


def table2d_init(nx, ny, values, dvdx=None, dvdy=None):
    """Returns coeff(nboxs,4,4) as a Fortran-ordered flat array."""
    ix1 = [0, 1, 1, 0]; ix2 = [0, 0, 1, 1]
    A = np.zeros((16, 16))
    for ic in range(4):
        n1, n2 = ix1[ic], ix2[ic]
        for p1 in range(4):
            for p2 in range(4):
                p1m, p2m = max(p1 - 1, 0), max(p2 - 1, 0)
                col = 4 * p1 + p2
                A[ic, col] = 1.0 * (n1 ** p1 * n2 ** p2)
                A[ic + 4, col] = 1.0 * (p1 * n1 ** p1m * n2 ** p2)
                A[ic + 8, col] = 1.0 * (n1 ** p1 * p2 * n2 ** p2m)
                A[ic + 12, col] = 1.0 * (p1 * n1 ** p1m * p2 * n2 ** p2m)
    nboxs = nx * ny
    B = np.zeros((16, nboxs))
    for nh in range(nx):
        for nc in range(ny):
            col = ny * nh + nc
            for ic in range(4):
                n1, n2 = ix1[ic] + nh, ix2[ic] + nc
                B[ic, col] = values[n1, n2]
                if dvdx is not None:
                    B[ic + 4, col] = dvdx[n1, n2]
                if dvdy is not None:
                    B[ic + 8, col] = dvdy[n1, n2]
    X = np.linalg.solve(A, B)
    coeff = np.zeros((nboxs, 4, 4))
    for i in range(4):
        for j in range(4):
            coeff[:, i, j] = X[4 * i + j, :]
    return np.asfortranarray(coeff).ravel(order='F').copy()


def table3d_init(nx, ny, nz, values, dvdx=None, dvdy=None, dvdz=None):
    ix1 = [0, 1, 1, 0, 0, 1, 1, 0]; ix2 = [0, 0, 1, 1, 0, 0, 1, 1]; ix3 = [0, 0, 0, 0, 1, 1, 1, 1]
    A = np.zeros((64, 64))
    for ic in range(8):
        n1, n2, n3 = ix1[ic], ix2[ic], ix3[ic]
        for p1 in range(4):
            for p2 in range(4):
                for p3 in range(4):
                    p1m, p2m, p3m = max(p1 - 1, 0), max(p2 - 1, 0), max(p3 - 1, 0)
                    col = 16 * p1 + 4 * p2 + p3
                    A[ic, col] = 1.0 * (n1 ** p1 * n2 ** p2 * n3 ** p3)
                    A[ic + 8, col] = 1.0 * (p1 * n1 ** p1m * n2 ** p2 * n3 ** p3)
                    A[ic + 16, col] = 1.0 * (n1 ** p1 * p2 * n2 ** p2m * n3 ** p3)
                    A[ic + 24, col] = 1.0 * (n1 ** p1 * n2 ** p2 * p3 * n3 ** p3m)
                    A[ic + 32, col] = 1.0 * (p1 * n1 ** p1m * p2 * n2 ** p2m * n3 ** p3)
                    A[ic + 40, col] = 1.0 * (p1 * n1 ** p1m * n2 ** p2 * p3 * n3 ** p3m)
                    A[ic + 48, col] = 1.0 * (n1 ** p1 * p2 * n2 ** p2m * p3 * n3 ** p3m)
                    A[ic + 56, col] = 1.0 * (p1 * n1 ** p1m * p2 * n2 ** p2m * p3 * n3 ** p3m)
    nboxs = nx * ny * nz
    B = np.zeros((64, nboxs))
    for ni in range(nx):
        for nj in range(ny):
            for nc in range(nz):
                col = nx * (ny * nc + nj) + ni
                for ic in range(8):
                    n1, n2, n3 = ix1[ic] + ni, ix2[ic] + nj, ix3[ic] + nc
                    B[ic, col] = values[n1, n2, n3]
                    if dvdx is not None:
                        B[ic + 8, col] = dvdx[n1, n2, n3]
                    if dvdy is not None:
                        B[ic + 16, col] = dvdy[n1, n2, n3]
                    if dvdz is not None:
                        B[ic + 24, col] = dvdz[n1, n2, n3]
    X = np.linalg.solve(A, B)
    coeff = np.zeros((nboxs, 4, 4, 4))
    for i in range(4):
        for j in range(4):
            for k in range(4):
                coeff[:, i, j, k] = X[16 * i + 4 * j + k, :]
    return np.asfortranarray(coeff).ravel(order='F').copy()


def make_cc_g_spline():
    """rebo2_db.f90:405-524; returns (g1_coeff, g2_coeff) as c(6,3) Fortran-flat arrays."""
    th = CC_G_THETA
    g1c = np.zeros((6, 3)); g2c = np.zeros((6, 3))
    A = np.zeros((6, 6))
    for i in range(3, 7):
        z = th[i - 1]
        for j in range(1, 7):
            A[i - 3, j - 1] = z ** (j - 1)
    z = th[2]
    A[4, 1] = 1.0
    A[5, 2] = 2.0
    for j in range(3, 7):
        A[4, j - 1] = (j - 1) * z ** (j - 2)
        if j >= 4:
            A[5, j - 1] = (j - 2) * (j - 1) * z ** (j - 3)
    B = np.array(CC_G_G1[2:6] + [CC_G_DG1[2], CC_G_D2G1[2]])
    g1c[:, 2] = np.linalg.solve(A, B)
    B = np.array(CC_G_G2[2:6] + [CC_G_DG1[2], CC_G_D2G1[2]])
    g2c[:, 2] = np.linalg.solve(A, B)
    for k in range(2):
        A = np.zeros((6, 6))
        for i in range(2):
            z = th[k] * (1 - i) + th[1 + k] * i
            A[3 * i, 0] = 1.0
            A[3 * i + 1, 1] = 1.0
            A[3 * i + 2, 2] = 2.0
            for j in range(2, 7):
                A[3 * i, j - 1] = z ** (j - 1)
                if j >= 3:
                    A[3 * i + 1, j - 1] = (j - 1) * z ** (j - 2)
                if j >= 4:
                    A[3 * i + 2, j - 1] = (j - 2) * (j - 1) * z ** (j - 3)
        B = np.array([CC_G_G1[k], CC_G_DG1[k], CC_G_D2G1[k], CC_G_G1[1 + k], CC_G_DG1[1 + k], CC_G_D2G1[1 + k]])
        x = np.linalg.solve(A, B)
        g1c[:, k] = x
        g2c[:, k] = x
    return g1c.ravel(order='F').copy(), g2c.ravel(order='F').copy()


class Rebo2:
    C_, H_ = 1, 3

    def __init__(self, **kwargs):
        d = dict(REBO2_DEFAULTS)
        tabs = rebo2_default_tables()
        for k, v in kwargs.items():
            if k in tabs:
                tabs[k] = np.asarray(v, dtype=np.float64)
            elif k in d:
                d[k] = v
            else:
                raise KeyError(k)
        self.d = d
        p = Rebo2Params()
        for k, _ in Rebo2Params._fields_[:19]:
            setattr(p, k, d[k])
        for i in range(6):
            p.cc_g_theta[i] = CC_G_THETA[i]
        g1c, g2c = make_cc_g_spline()
        for i in range(18):
            p.cc_g1_coeff[i] = g1c[i]
            p.cc_g2_coeff[i] = g2c[i]
            p.spgh[i] = SPGH[i]
        for i in range(25):
            p.igh[i] = IGH[i]
        # rebo2_db.f90:147-168
        for t in (0, 2):
            p.conpe[t] = -0.5
            p.conan[t] = 0.5 * -0.5
            p.conpf[t] = -0.5 - 1.0
        p.conalp = d['hhh_lambda']
        ce = np.zeros((6, 6))
        CC, CH, HH = 0, 2, 5
        al = d['hhh_lambda']
        ce[CC, CC] = 1.0
        ce[CC, CH] = np.exp(al * (d['ch_re'] - d['cc_re']))
        ce[CC, HH] = np.exp(al * (d['hh_re'] - d['cc_re']))
        ce[CH, CC] = 1.0 / ce[CC, CH]
        ce[CH, CH] = 1.0
        ce[CH, HH] = np.exp(al * (d['hh_re'] - d['ch_re']))
        ce[HH, CC] = 1.0 / ce[CC, HH]
        ce[HH, CH] = 1.0 / ce[CH, HH]
        ce[HH, HH] = 1.0
        flat = ce.ravel(order='F')
        for i in range(36):
            p.conear[i] = flat[i]
        for idx, (l, h) in {0: ('cc_in_r1', 'cc_in_r2'), 2: ('ch_r1', 'ch_r2'), 5: ('hh_r1', 'hh_r2')}.items():
            p.cut_in_l[idx] = d[l]
            p.cut_in_h[idx] = d[h]
            p.cut_in_h2[idx] = d[h] ** 2
        p.with_dihedral = int(bool(d['with_dihedral']))
        self._keep = {}
        self._keep['Fcc'] = table3d_init(4, 4, 9, tabs['Fcc'], tabs['dFdi'], tabs['dFdj'], tabs['dFdk'])
        self._keep['Fch'] = table3d_init(4, 4, 9, tabs['Fch'])
        self._keep['Fhh'] = table3d_init(4, 4, 9, tabs['Fhh'])
        self._keep['Tcc'] = table3d_init(4, 4, 9, tabs['Tcc'])
        self._keep['Pcc'] = table2d_init(5, 5, tabs['Pcc'])
        self._keep['Pch'] = table2d_init(5, 5, tabs['Pch'])
        for k in ('Fcc', 'Fch', 'Fhh', 'Tcc'):
            t = getattr(p, k)
            t.nx, t.ny, t.nz, t.coeff = 4, 4, 9, _p(self._keep[k])
        for k in ('Pcc', 'Pch'):
            t = getattr(p, k)
            t.nx, t.ny, t.coeff = 5, 5, _p(self._keep[k])
        self.p = p
        self.tabs = tabs

    def cutoff(self, symbols):
        s = set(symbols)
        c = 0.0
        if 'C' in s:
            c = max(c, self.d['cc_in_r2'])
        if 'C' in s and 'H' in s:
            c = max(c, self.d['ch_r2'])
        if 'H' in s:
            c = max(c, self.d['hh_r2'])
        return c

    def ktyp(self, symbols):
        return np.array([1 if s == 'C' else 3 if s == 'H' else 0 for s in symbols], dtype=np.int32)

    def table2d_eval(self, name, a, b):
        v = (C.c_double * 3)()
        lib().orc_table2d_eval(C.byref(getattr(self.p, name)), C.c_double(a), C.c_double(b),
                               C.byref(v, 0), C.byref(v, 8), C.byref(v, 16))
        return tuple(v)

    def table3d_eval(self, name, a, b, c):
        v = (C.c_double * 4)()
        lib().orc_table3d_eval(C.byref(getattr(self.p, name)), C.c_double(a), C.c_double(b), C.c_double(c),
                               C.byref(v, 0), C.byref(v, 8), C.byref(v, 16), C.byref(v, 24))
        return tuple(v)

    def energy_and_forces(self, r, cell, nl, ktyp, per_at=False, per_bond=False):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nat = len(r)
        abox = abox_from_cell(cell)
        ktyp = np.ascontiguousarray(ktyp, dtype=np.int32)
        epot = C.c_double(0.0)
        f = np.zeros((nat, 3))
        wpot = np.zeros(9)
        epa = np.zeros(nat) if per_at else None
        wpa = np.zeros((nat, 9)) if per_at else None
        epb, fpb, wpb = _per_bond_arrays(nl, per_bond)
        err = lib().orc_rebo2_energy_and_forces(
            C.byref(self.p), C.c_int(nat), C.c_int(nat), _p(r), _p(abox), _p(ktyp, C.c_int),
            _p(nl.seed, C.c_ssize_t), _p(nl.last, C.c_ssize_t), _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int),
            C.byref(epot), _p(f), _p(wpot), _p(epa), _p(epb), _p(fpb), _p(wpa), _p(wpb))
        assert err == 0
        out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
        if per_at:
            out['epot_per_at'] = epa
            out['wpot_per_at'] = wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy()
        if per_bond:
            out['epot_per_bond'] = epb
            out['f_per_bond'] = fpb
            out['wpot_per_bond'] = wpb.reshape(-1, 3, 3).transpose(0, 2, 1).copy()
        return out


class Rebo2ScrStruct(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('cc_ar_r1', 'cc_ar_r2', 'cc_bo_r1', 'cc_bo_r2', 'cc_nc_r1', 'cc_nc_r2',
                                          'Cmin', 'Cmax')]


# rebo2_type.f90:65-66, 204-213 (SCREENING branch)
REBO2_SCR_DEFAULTS = dict(cc_in_r1=1.95, cc_in_r2=2.25, cc_ar_r1=2.179347, cc_ar_r2=2.819732,
                          cc_bo_r1=1.866344, cc_bo_r2=2.758372, cc_nc_r1=1.217335, cc_nc_r2=4.000000,
                          Cmin=1.00, Cmax=2.00)


class Rebo2Scr(Rebo2):
    """rebo2_scr (src/potentials/bop/rebo2/rebo2_scr.f90): screened REBO2, C-C bonds screened only,
    trigonometric cutoffs; with_dihedral=True adds the alternative dihedral term (ALT_DIHEDRAL,
    bop_kernel_rebo2.f90:2089-2371)"""

    def __init__(self, **kwargs):
        sd = dict(REBO2_SCR_DEFAULTS)
        base = {}
        for k, v in kwargs.items():
            if k in sd:
                sd[k] = v
            else:
                base[k] = v
        base.setdefault('cc_in_r1', sd['cc_in_r1'])
        base.setdefault('cc_in_r2', sd['cc_in_r2'])
        super().__init__(**base)
        self.sd = sd
        q = Rebo2ScrStruct()
        for k, _ in Rebo2ScrStruct._fields_:
            setattr(q, k, sd[k])
        self.q = q

    def cutoff(self, symbols):
        # rebo2_module.f90:95-119
        cmax = self.sd['Cmax']
        c_cc = np.sqrt(cmax * cmax / (4 * (cmax - 1))) * max(self.d['cc_in_r2'], self.sd['cc_ar_r2'],
                                                             self.sd['cc_bo_r2'], self.sd['cc_nc_r2'])
        s = set(symbols)
        c = 0.0
        if 'C' in s:
            c = max(c, c_cc)
        if 'H' in s:
            c = max(c, self.d['hh_r2'])
        return c

    def energy_and_forces(self, r, cell, nl, ktyp, per_at=False, per_bond=False):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nat = len(r)
        abox = abox_from_cell(cell)
        ktyp = np.ascontiguousarray(ktyp, dtype=np.int32)
        epot = C.c_double(0.0)
        f = np.zeros((nat, 3))
        wpot = np.zeros(9)
        epa = np.zeros(nat) if per_at else None
        wpa = np.zeros((nat, 9)) if per_at else None
        epb, fpb, wpb = _per_bond_arrays(nl, per_bond)
        err = lib().orc_rebo2_scr_energy_and_forces(
            C.byref(self.p), C.byref(self.q), C.c_int(nat), C.c_int(nat), _p(r), _p(abox), _p(ktyp, C.c_int),
            _p(nl.seed, C.c_ssize_t), _p(nl.last, C.c_ssize_t), _p(nl.neighbors, C.c_int), _p(nl.dc, C.c_int),
            C.byref(epot), _p(f), _p(wpot), _p(epa), _p(epb), _p(fpb), _p(wpa), _p(wpb))
        assert err == 0, err
        out = dict(epot=epot.value, f=f, wpot=wpot.reshape(3, 3).T.copy())
        if per_at:
            out['epot_per_at'] = epa
            out['wpot_per_at'] = wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy()
        if per_bond:
            out['epot_per_bond'] = epb
            out['f_per_bond'] = fpb
            out['wpot_per_bond'] = wpb.reshape(-1, 3, 3).transpose(0, 2, 1).copy()
        return out


BOP_FUNCS = dict(VA=0, VR=1, g=2, bo=3, h=4)


def bop_func(params, which, x, ktypj=1, ktypi=1, ktypk=1, ijpot=1, ikpot=1, fcij=1.0, faij=1.0):
    """one evaluation of VA / VR / g / bo / h of the *_func.f90 files (1-based indices): (value, derivative);
    for bo: (bij, dfbij)"""
    v, d = C.c_double(0.0), C.c_double(0.0)
    lib().orc_bop_func(C.byref(params), C.c_int(BOP_FUNCS[which]), C.c_int(ktypj), C.c_int(ktypi), C.c_int(ktypk),
                       C.c_int(ijpot), C.c_int(ikpot), C.c_double(x), C.c_double(fcij), C.c_double(faij),
                       C.byref(v), C.byref(d))
    return v.value, d.value


REBO2_FUNCS = dict(fconj=0, fCin=1, VA=2, VR=3, g=4, bo=5, h=6, Z2pair=7)


def rebo2_func(params, which, x=0.0, y=0.0, z=0.0, i1=1, i2=1):
    """one evaluation of a function of rebo2_func.f90: the three result slots of orc_rebo2_func"""
    out = (C.c_double * 3)()
    lib().orc_rebo2_func(C.byref(params), C.c_int(REBO2_FUNCS[which]), C.c_int(i1), C.c_int(i2), C.c_double(x),
                         C.c_double(y), C.c_double(z), out)
    return tuple(out)


def cutoff_eval(kind, r1, r2, r):
    """kind 'trig_off' or 'exp': (value, derivative) of the cutoff function between r1 and r2"""
    v, d = C.c_double(0.0), C.c_double(0.0)
    lib().orc_cutoff_eval(C.c_int(0 if kind == 'trig_off' else 1), C.c_double(r1), C.c_double(r2), C.c_double(r),
                          C.byref(v), C.byref(d))
    return v.value, d.value
