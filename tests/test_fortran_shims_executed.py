"""The hand-written Fortran shims of atomistica_b200/fortran/ EXECUTED (no Fortran compiler exists in this image):
tests/fortran_subset.py translates them like the reference's kernels, the C ABI is replaced by recorders, and the
parameter structs the shims fill from the reference's derived types are compared, component by component, with
what the Python host (atomistica_b200.native) hands to the same entry points -- which is what the GPU tests run."""
import os
import types

import numpy as np
import pytest

from atomistica_b200 import _lib as L, native, parameters as P
from atomistica_b200.elements import atomic_numbers
from fortran_subset import FA, Obj, units

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FDIR = os.path.join(ROOT, 'atomistica_b200', 'fortran')


def _types():
    from test_abi import _fortran_types
    return _fortran_types(open(os.path.join(FDIR, 'atx_c_api.f90')).read())


def _new_type(name):
    """an instance of a bind(C) derived type of atx_c_api.f90 (arrays zeroed, pointers null)"""
    o = Obj()
    for comp, ftype, dim in _types()[name]:
        setattr(o, comp if comp != 'lambda' else 'lambda_', FA(dim) if dim else (None if ftype == 'type(c_ptr)' else 0))
    return o


def _db_object(db, el_chars=True):
    nel = len(db['el'])
    fields = {('lambda_' if k == 'lambda' else k): FA(len(v), data=list(v)) for k, v in db.items()
              if k not in ('__ref__', 'el')}
    chars = []
    for s in db['el']:
        chars += list(s.ljust(2))
    return Obj(nel=nel, el=FA(2, nel, data=chars), **fields)


class Recorder:
    def __init__(self):
        self.calls = []

    def fn(self, name):
        def call(*args):
            self.calls.append((name, args))
            return 0
        return call


KINDS = dict(Tersoff=1, Kumagai=2, Brenner=3)


@pytest.mark.parametrize('kind,dbname,screened', [
    ('Tersoff', 'Tersoff_PRB_39_5566_Si_C', False), ('Tersoff', 'Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N', False),
    ('Kumagai', 'Kumagai_CompMaterSci_39_457_Si', False), ('Brenner', 'Erhart_PRB_71_035211_SiC', False),
    ('Tersoff', 'Tersoff_PRB_39_5566_Si_C__Scr', True), ('Kumagai', 'Kumagai_CompMaterSci_39_457_Si__Scr', True),
    ('Brenner', 'Erhart_PRB_71_035211_SiC__Scr', True)])
def test_bop_bind_to_shim_fills_the_parameter_structs_like_the_python_host(kind, dbname, screened):
    db = (P.complete_scr if screened else P.complete)(kind, getattr(P, dbname))
    rec = Recorder()
    macros = {'BOP_TYPE': (None, 'bop_t'), 'ATX_BOP_KIND': (None, str(KINDS[kind])), 'BOP_NAME_STR': (None, '"x"'),
              'COMPUTE_FUNC': (None, 'compute_func')}
    env = dict(atx_ctx='ctx', _new_type=_new_type, a2s=lambda fa: ''.join(fa.data).strip(),
               atomic_number=lambda s: atomic_numbers.get(s, 0), **{'ATX_BOP_%s' % k.upper(): v for k, v in KINDS.items()})
    for name in ('atx_bop_create', 'atx_bop_create_screened', 'atx_bop_bind_to', 'atx_bop_destroy'):
        env[name] = rec.fn(name)
    fn = units(open(os.path.join(FDIR, 'bop_compute_gpu.f90')).read(), defined={'SCREENING'} if screened else set(),
               env=env, macros=macros, noops=('atx_pass_error', 'timer_start', 'timer_stop', 'update'))['bop_bind_to_gpu']
    assert callable(fn), fn
    this = Obj(db=_db_object(db), atx_pot=None)
    p = Obj(nel=len(db['el']), el2Z=FA(len(db['el']), data=[atomic_numbers[s] for s in db['el']]))
    fn(this, p, Obj(atx_p='p', atx_nl='nl'))
    names = [c[0] for c in rec.calls]
    assert names == (['atx_bop_create_screened'] if screened else ['atx_bop_create']) + ['atx_bop_bind_to'], names
    create = rec.calls[0][1]
    assert create[0] == 'ctx'
    par = create[1]
    want = native._Bop._fill(types.SimpleNamespace(kind=kind), db)
    for cname, ctype in L.AtxBopParams._fields_:
        got = getattr(par, cname)
        ref = getattr(want, cname)
        if hasattr(ctype, '_length_'):
            assert list(got) == list(ref), (kind, cname, list(got), list(ref))
        else:
            assert got == ref, (kind, cname)
    if screened:
        scr = create[2]
        npairs = len(db['el']) * (len(db['el']) + 1) // 2
        for key in P.SCR_KEYS:
            assert list(getattr(scr, key))[:npairs] == [float(x) for x in db[key][:npairs]], key
    bind = rec.calls[-1][1]
    assert bind[1:3] == ('p', 'nl') and bind[3] == len(db['el']) and list(bind[4]) == [atomic_numbers[s] for s in db['el']]


@pytest.mark.parametrize('screened', [False, True])
def test_rebo2_bind_to_shim(screened):
    """rebo2_gpu.f90: the rebo2_t image carries the constants the REFERENCE's own statements produce (rebo2_db.f90, run
    through the translator as in tests/test_func_vs_reference.py), the shim copies them into atx_rebo2_params_t, and the
    result equals what rebo2_tables.build_params hands to atx_rebo2_create from Python"""
    import oracle
    import test_func_vs_reference as T
    from atomistica_b200 import rebo2_tables
    if not os.path.isdir(T.BOP):
        pytest.skip('the reference tree is not present')
    orc = (oracle.Rebo2Scr if screened else oracle.Rebo2)(with_dihedral=True)
    this = T._rebo2_this(orc)
    this.with_dihedral = True
    this.atx_pot = None
    tabs = oracle.rebo2_default_tables()
    for name, extra in (('Fcc', ('dFdi', 'dFdj', 'dFdk')), ('Fch', ()), ('Fhh', ()), ('Tcc', ())):
        c = oracle.table3d_init(4, 4, 9, tabs[name], *[tabs[k] for k in extra])
        setattr(this, name, Obj(coeff=FA(144, 4, 4, 4, data=list(c))))
    for name in ('Pcc', 'Pch'):
        setattr(this, name, Obj(coeff=FA(25, 4, 4, data=list(oracle.table2d_init(5, 5, tabs[name])))))
    if screened:
        this.__dict__.update(orc.sd)
    rec = Recorder()
    env = dict(atx_ctx='ctx', _new_type=_new_type)
    for name in ('atx_rebo2_create', 'atx_rebo2_create_screened', 'atx_rebo2_bind_to', 'atx_rebo2_destroy'):
        env[name] = rec.fn(name)
    fn = units(open(os.path.join(FDIR, 'rebo2_gpu.f90')).read(), defined={'SCREENING'} if screened else set(), env=env,
               macros={'BOP_TYPE': (None, 'rebo2_t'), 'COMPUTE_FUNC': (None, 'compute_func'), 'BOP_NAME_STR': (None, '"r"')},
               noops=('atx_pass_error', 'timer_start', 'timer_stop', 'update'))['rebo2_bind_to_gpu']
    assert callable(fn), fn
    fn(this, Obj(nel=2, el2Z=FA(2, data=[6, 1])), Obj(atx_p='p', atx_nl='nl'))
    assert [c[0] for c in rec.calls] == ['atx_rebo2_create_screened' if screened else 'atx_rebo2_create', 'atx_rebo2_bind_to']
    par = rec.calls[0][1][1]
    kw = dict(with_dihedral=True)
    if screened:
        kw.update(cc_in_r1=orc.d['cc_in_r1'], cc_in_r2=orc.d['cc_in_r2'])
    _, _, want, keep = rebo2_tables.build_params(kw)
    for cname, ctype in L.AtxRebo2Params._fields_:
        got, ref = getattr(par, cname), getattr(want, cname)
        if cname in keep:                                   # table pointers: c_loc of the first coefficient
            assert abs(got - keep[cname][0]) <= 1e-12 * max(1.0, abs(keep[cname][0])), cname
        elif hasattr(ctype, '_length_'):
            got, ref = np.asarray(list(got), float), np.asarray(list(ref), float)
            tol = 1e-9 * max(1.0, np.abs(ref).max()) if 'coeff' in cname else 0.0      # two different linear solvers
            assert np.abs(got - ref).max() <= tol, (cname, np.abs(got - ref).max())
        else:
            assert got == ref, (cname, got, ref)
    if screened:
        scr = rec.calls[0][1][2]
        for key in ('cc_ar_r1', 'cc_ar_r2', 'cc_bo_r1', 'cc_bo_r2', 'cc_nc_r1', 'cc_nc_r2', 'Cmin', 'Cmax'):
            assert getattr(scr, key) == orc.sd[key], key


def _prototypes():
    import importlib.util
    spec = importlib.util.spec_from_file_location('gen_fortran_api', os.path.join(ROOT, 'scripts', 'gen_fortran_api.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = {}
    for ret, name, args in gen.prototypes(open(gen.HEADER).read()):
        out[name] = [a.replace('*', ' ').split()[-1] for a in args]
    return out


class Tagged(FA):
    """an array whose elements say which array and which element they are: c_loc(x(1, 1)) is then recognisable"""

    def __init__(self, tag, *shape):
        n = int(np.prod(shape))
        FA.__init__(self, *shape, data=['%s[%d]' % (tag, k) for k in range(n)])
        self.tag = tag


@pytest.mark.parametrize('shim,unit,entry,optionals', [
    ('bop_compute_gpu.f90', 'compute_func', 'atx_bop_energy_and_forces',
     ('mask', 'epot_per_at', 'epot_per_bond', 'f_per_bond', 'wpot_per_at', 'wpot_per_bond')),
    ('rebo2_gpu.f90', 'compute_func', 'atx_rebo2_energy_and_forces',
     ('epot_per_at', 'epot_per_bond', 'f_per_bond', 'wpot_per_at', 'wpot_per_bond')),
    ('tabulated_alloy_eam_gpu.f90', 'tabulated_alloy_eam_energy_and_forces', 'atx_eam_energy_and_forces',
     ('mask', 'epot_per_at', 'wpot_per_at'))])
def test_compute_shims_pass_every_argument_in_the_position_the_header_declares(shim, unit, entry, optionals):
    """COMPUTE_FUNC of the three families: absent optionals arrive as NULL, present ones as the address of their first
    element, every argument in the position of the C prototype's parameter of the same name"""
    rec = Recorder()
    fn = units(open(os.path.join(FDIR, shim)).read(), env={entry: rec.fn(entry), 'ATX_BOP_KIND': 1},
               macros={'BOP_TYPE': (None, 'pot_t'), 'COMPUTE_FUNC': (None, 'compute_func'), 'BOP_NAME_STR': (None, '"x"'),
                       'ATX_BOP_KIND': (None, '1')},
               noops=('atx_pass_error', 'timer_start', 'timer_stop', 'update'))[unit]
    assert callable(fn), fn
    params = _prototypes()[entry]
    this, p, nl = Obj(atx_pot='POT'), Obj(maxnatloc=4, nat=4), Obj(atx_p='P', atx_nl='NL', neighbors_size=9)
    shapes = dict(mask=(4,), epot_per_at=(4,), epot_per_bond=(9,), f_per_bond=(3, 9), wpot_per_at=(3, 3, 4), wpot_per_bond=(3, 3, 9))
    f, wpot = Tagged('f', 3, 4), Tagged('wpot', 3, 3)
    for present in (False, True):
        rec.calls.clear()
        kw = {k: Tagged(k, *shapes[k]) for k in optionals} if present else {}
        fn(this, p, nl, 1.5, f, wpot, **kw)
        (name, args), = rec.calls
        assert len(args) == len(params), (len(args), params)
        for value, pname in zip(args, params):
            if pname == 'pot':
                assert value == 'POT'
            elif pname == 'p':
                assert value == 'P'
            elif pname == 'nl':
                assert value == 'NL'
            elif pname == 'epot':
                assert value == 1.5
            elif pname in ('f', 'wpot'):
                assert value is (f if pname == 'f' else wpot)
            else:
                assert pname in optionals, pname
                assert value == ('%s[0]' % pname if present else None), (pname, value)


def test_eam_bind_to_shim_hands_over_the_spline_arrays():
    """tabulated_alloy_eam_gpu.f90: one atx_spline_t image per simple_spline_t (F, rho per element, phi per pair), each
    pointing at the first element of the host arrays; then atx_eam_bind_to with el2db"""
    rec = Recorder()
    env = dict(atx_ctx='ctx', _new_type=_new_type)
    for name in ('atx_eam_create', 'atx_eam_bind_to', 'atx_eam_destroy'):
        env[name] = rec.fn(name)
    # automatic arrays of derived type: type(atx_spline_t) :: fF(n) -- elements created on first use
    fn = units(open(os.path.join(FDIR, 'tabulated_alloy_eam_gpu.f90')).read(), env=env,
               noops=('atx_pass_error', 'timer_start', 'timer_stop', 'update'))
    assert callable(fn['tabulated_alloy_eam_bind_to_gpu']), fn['tabulated_alloy_eam_bind_to_gpu']
    assert callable(fn['spline_image']), fn['spline_image']

    def spline(tag, n):
        return Obj(n=n, x0=0.0, dx=0.01 * n, **{k: Tagged('%s.%s' % (tag, k), n if k == 'y' else n - 1)
                                                for k in ('y', 'coeff1', 'coeff2', 'coeff3', 'dcoeff1', 'dcoeff2', 'dcoeff3')})
    this = Obj(db=Obj(nel=2), atx_pot=None, cutoff=5.5, el2db=FA(3, data=[2, -1, 1]),
               fF=FA(2, data=[spline('F1', 11), spline('F2', 12)]), frho=FA(2, data=[spline('r1', 21), spline('r2', 22)]),
               fphi=FA(2, 2, data=[spline('p11', 31), spline('p21', 32), spline('p12', 33), spline('p22', 34)]))
    fn['tabulated_alloy_eam_bind_to_gpu'](this, Obj(nel=3), Obj(atx_p='P', atx_nl='NL'))
    assert [c[0] for c in rec.calls] == ['atx_eam_create', 'atx_eam_bind_to']
    ctx, ndb, fF, frho, fphi, cutoff, _ = rec.calls[0][1]
    assert (ctx, ndb, cutoff) == ('ctx', 2, 5.5)
    assert fF(2).n == 12 and fF(2).y == 'F2.y[0]' and fF(1).dcoeff3 == 'F1.dcoeff3[0]' and fF(1).dx == 0.11
    assert frho(1).coeff1 == 'r1.coeff1[0]'
    assert fphi(2, 1).y == 'p21.y[0]' and fphi(1, 2).coeff2 == 'p12.coeff2[0]'        # column-major (i, j)
    pot, P_, NL, nel, el2db = rec.calls[1][1]
    assert (P_, NL, nel) == ('P', 'NL', 3) and list(el2db) == [2, -1, 1]


def _abi(rec, name, outs=(), **results):
    """a recorder for one C entry point seen from Fortran: positional dummies from the header, the named ones returned"""
    params = _prototypes()[name]

    def call(*args):
        rec.calls.append((name, args))
        return dict(result=0, **results)
    call.fortran_args = (tuple(params), tuple(outs))
    return call


def test_neighbour_list_shim_call_sequence():
    """python_neighbors_gpu.f90: first call creates context, list and device particles and requests the range; every
    call mirrors cell / elements / positions, updates, copies the list back in the reference's layout, reads the
    statistics"""
    rec = Recorder()
    env = dict(atx_ctx=None)
    env['atx_ctx_create'] = _abi(rec, 'atx_ctx_create', ('ctx',), ctx='CTX')
    env['atx_neighbors_create'] = _abi(rec, 'atx_neighbors_create', ('nl',), nl='NL')
    env['atx_particles_create'] = _abi(rec, 'atx_particles_create', ('p',), p='P')
    env['atx_neighbors_get_info'] = _abi(rec, 'atx_neighbors_get_info', ('npairs', 'nebmax'), npairs=1200, nebmax=7)
    for name in ('atx_neighbors_request_interaction_range', 'atx_particles_set_cell', 'atx_particles_set_elements',
                 'atx_particles_set_positions', 'atx_neighbors_update', 'atx_neighbors_copy_to_host'):
        env[name] = rec.fn(name)
    fn = units(open(os.path.join(FDIR, 'python_neighbors_gpu.f90')).read(), env=env, global_scalars=('atx_ctx',),
               noops=('atx_pass_error', 'timer_start', 'timer_stop'))['fill_neighbor_list']
    assert callable(fn), fn
    this = Obj(atx_nl=None, atx_p=None, avgn=100, cutoff=5.5, seed=Tagged('seed', 5), last=Tagged('last', 5),
               neighbors=Tagged('neighbors', 400), dc=Tagged('dc', 3, 400), neighbors_size=400, nupdate=0, avgnn=0.0)
    p = Obj(nat=4, Abox=Tagged('Abox', 3, 3), Bbox=Tagged('Bbox', 3, 3), pbc=Tagged('pbc', 3), el=Tagged('el', 4),
            r_non_cyc=Tagged('r', 3, 4))
    fn(this, p)
    names = [c[0] for c in rec.calls]
    assert names == ['atx_ctx_create', 'atx_neighbors_create', 'atx_particles_create', 'atx_neighbors_request_interaction_range',
                     'atx_particles_set_cell', 'atx_particles_set_elements', 'atx_particles_set_positions',
                     'atx_neighbors_update', 'atx_neighbors_copy_to_host', 'atx_neighbors_get_info']
    calls = dict(rec.calls)
    assert (this.atx_nl, this.atx_p) == ('NL', 'P')
    assert calls['atx_neighbors_create'][:2] == ('CTX', 100) and calls['atx_particles_create'][0] == 'CTX'
    assert calls['atx_neighbors_request_interaction_range'] == ('NL', 5.5)
    assert calls['atx_particles_set_cell'][0] == 'P' and calls['atx_particles_set_cell'][1] is p.Abox \
        and calls['atx_particles_set_cell'][2] is p.Bbox and calls['atx_particles_set_cell'][3] is p.pbc
    assert calls['atx_particles_set_elements'][:2] == ('P', 4) and calls['atx_particles_set_elements'][2] is p.el
    assert calls['atx_particles_set_positions'][1] == 4 and calls['atx_particles_set_positions'][2] is p.r_non_cyc
    assert calls['atx_neighbors_update'] == ('NL', 'P')
    c = calls['atx_neighbors_copy_to_host']
    assert c[0] == 'NL' and c[1] is this.seed and c[2] is this.last and c[3] is this.neighbors and c[4] is this.dc and c[5] == 400
    assert this.nupdate == 1 and this.avgnn == 1200 / 4
    rec.calls.clear()
    fn(this, p)                                      # second call: nothing is created again
    assert [c[0] for c in rec.calls][:2] == ['atx_particles_set_cell', 'atx_particles_set_elements']


def test_particles_shims():
    rec = Recorder()
    env = {name: rec.fn(name) for name in ('atx_particles_set_cell', 'atx_particles_set_elements', 'atx_particles_set_positions')}
    fns = units(open(os.path.join(FDIR, 'python_particles_gpu.f90')).read(), env=env, noops=('atx_pass_error',))
    this = Obj(nat=3, Abox=Tagged('A', 3, 3), Bbox=Tagged('B', 3, 3), pbc=FA(3, data=[1, 0, 1]), el=Tagged('el', 3),
               r_non_cyc=Tagged('r', 3, 3))
    fns['particles_set_cell_gpu'](this, 'P')
    fns['particles_update_elements_gpu'](this, 'P')
    fns['particles_sync_positions_gpu'](this, 'P')
    (n1, a1), (n2, a2), (n3, a3) = rec.calls
    assert n1 == 'atx_particles_set_cell' and a1[0] == 'P' and a1[1] is this.Abox and a1[2] is this.Bbox and list(a1[3]) == [1, 0, 1]
    assert n2 == 'atx_particles_set_elements' and a2[:2] == ('P', 3) and a2[2] is this.el
    assert n3 == 'atx_particles_set_positions' and a3[:2] == ('P', 3) and a3[2] is this.r_non_cyc
