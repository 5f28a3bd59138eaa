"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/*.h declares, its
host-side init helpers agree with the oracle's independent numpy implementation, and it fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from atomistica_b200 import _lib as L, io, parameters as P
from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'atomistica_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(atx_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    decl = _declared_symbols()
    assert len(decl) >= 40
    for s in decl:
        assert hasattr(lib, s), 'missing symbol ' + s
    assert sorted(L.SYMBOLS) == decl


def test_no_cpu_fallback():
    import subprocess
    have_gpu = subprocess.run(['nvidia-smi', '-L'], capture_output=True).returncode == 0 \
        if os.path.exists('/usr/bin/nvidia-smi') else False
    if have_gpu:
        pytest.skip('a GPU is present')
    h = C.c_void_p()
    err = L.lib().atx_ctx_create(0, C.byref(h))
    assert err != 0
    assert 'no CPU fallback' in L.last_error()
    from atomistica_b200 import native
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        native.Particles()


def test_host_spline_init_matches_oracle(cu_setfl):
    y = cu_setfl['rho'][0]
    n, dx = len(y), float(cu_setfl['dr'])
    arrs = [np.zeros(n), np.zeros(n)] + [np.zeros(n - 1) for _ in range(6)]
    L.check(L.lib().atx_host_spline_init(n, C.c_double(0.0), C.c_double(dx), L.dptr(L.as_f64(y)),
                                         *[L.dptr(a) for a in arrs]))
    s = oracle.spline_init(n, 0.0, dx, y)
    for a, k in zip(arrs, ('y', 'd2y', 'c1', 'c2', 'c3', 'd1', 'd2', 'd3')):
        assert np.allclose(a, s[k], rtol=1e-13, atol=1e-300), k


def test_host_gaussn():
    rng = np.random.RandomState(0)
    A = rng.rand(7, 7) + 3 * np.eye(7)
    B = rng.rand(7, 4)
    Af = np.asfortranarray(A).ravel(order='F').copy()
    Bf = np.asfortranarray(B).ravel(order='F').copy()
    L.check(L.lib().atx_host_gaussn(7, L.dptr(Af), 4, L.dptr(Bf)))
    assert np.allclose(Bf.reshape(4, 7).T, np.linalg.solve(A, B), rtol=1e-12)
    Z = np.zeros(9)
    assert L.lib().atx_host_gaussn(3, L.dptr(Z), 1, L.dptr(np.ones(3))) != 0
    assert 'singular' in L.last_error()


def test_host_tables_match_oracle():
    from atomistica_b200 import rebo2_tables as T
    d, tabs, p, keep = T.build_params({})
    rb = oracle.Rebo2()
    for k in ('Fcc', 'Fch', 'Fhh', 'Tcc', 'Pcc', 'Pch'):
        assert np.allclose(keep[k], rb._keep[k], rtol=1e-10, atol=1e-13), k
    assert np.allclose(np.array(p.cc_g1_coeff), np.array(rb.p.cc_g1_coeff), rtol=1e-9, atol=1e-12)
    assert np.allclose(np.array(p.cc_g2_coeff), np.array(rb.p.cc_g2_coeff), rtol=1e-9, atol=1e-12)
    assert np.allclose(np.array(p.conear), np.array(rb.p.conear))
    for k in tabs:
        assert np.array_equal(tabs[k], rb.tabs[k]), k


def test_setfl_roundtrip(tmp_path, cu_setfl):
    fn = str(tmp_path / 'Cu.eam.alloy')
    io.write_setfl(fn, cu_setfl)
    t = io.read_setfl(fn)
    assert t['names'] == ['Cu'] and t['nF'] == 10001 and t['nr'] == 10001
    assert t['cutoff'] == float(cu_setfl['cutoff'])
    for k in ('F', 'rho', 'rphi'):
        assert np.array_equal(t[k], cu_setfl[k])


def test_parameter_completion():
    db = P.complete('Tersoff', P.Goumri_Said_ChemPhys_302_135_Al_N)
    assert db['omega'] == [1.0, 1.0, 1.0] and db['mubo'] == [0.0, 0.0, 0.0] and db['m'] == [1, 1, 1]
    assert P.complete('Kumagai', None)['r2'] == [3.30]
    assert P.pair_index(0, 1, 2) == 1 and P.pair_index(1, 1, 2) == 2 and P.pair_index(2, 1, 3) == 4
    m = P.Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N
    assert abs(m['A'][1] - np.sqrt(1.3936e3 * 1.1e4)) < 1e-9


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """the drop-in boundary is a C header: it must compile as C99 (Fortran/C callers) and as C++, and
    the ctypes mirrors in atomistica_b200/_lib.py must have the compiler's struct sizes"""
    import subprocess
    inc = os.path.join(ROOT, 'include')
    structs = dict(atx_spline=L.AtxSpline, atx_bop_params=L.AtxBopParams, atx_bop_screening=L.AtxBopScreening,
                   atx_juslin_params=L.AtxJuslinParams, atx_pair_params=L.AtxPairParams,
                   atx_rebo2_params=L.AtxRebo2Params, atx_rebo2_screening=L.AtxRebo2Screening)
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include "atomistica_b200.h"\nint main(void) {\n' +
                   ''.join('  printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in structs) +
                   '  return 0;\n}\n')
    exe = tmp_path / 'sizes'
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I' + inc, str(src), '-o', str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes = dict(zip(out[::2], map(int, out[1::2])))
    for name, cls in structs.items():
        assert sizes[name] == C.sizeof(cls), (name, sizes[name], C.sizeof(cls))
    subprocess.run(['g++', '-std=c++17', '-pedantic', '-Wall', '-Werror', '-I' + inc, '-x', 'c++', '-c', str(src),
                    '-o', str(tmp_path / 'sizes.o')], check=True)


def _build_c_example(tmp_path):
    import subprocess
    libdir = os.path.join(ROOT, 'atomistica_b200')
    exe = str(tmp_path / 'tersoff_from_c')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I' + os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'examples', 'tersoff_from_c.c'), '-o', exe, '-L' + libdir,
                    '-latomistica_b200', '-Wl,-rpath,' + libdir, '-lm'], check=True)
    return exe


def test_c_example_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/tersoff_from_c.c is the C ABI used from plain C; without a GPU it must stop at
    atx_ctx_create with the library's error text"""
    import subprocess
    exe = _build_c_example(tmp_path)
    if os.path.exists('/usr/bin/nvidia-smi') and subprocess.run(['nvidia-smi', '-L'],
                                                                capture_output=True).returncode == 0:
        pytest.skip('a GPU is present')
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2
    assert 'no CPU fallback' in r.stderr


@pytest.mark.gpu
def test_c_example_on_the_gpu(tmp_path):
    import subprocess
    r = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    e0 = float(lines[1].split('=')[1].split()[0])
    assert abs(e0 - (-4.6295950127)) < 1e-8          # closed form, tests/golden/kat.json
    assert 'max |f|' in lines[2]
    fsum = [abs(float(x)) for x in lines[2].split('sum f =')[1].split()]
    assert max(fsum) < 1e-9


def test_fortran_interfaces_cover_the_whole_c_abi():
    """atomistica_b200/fortran/atx_c_api.f90 is generated from the header (scripts/gen_fortran_api.py):
    it is current, and every exported symbol has a bind(C) interface with as many dummy arguments as the C
    prototype has parameters"""
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location('gen_fortran_api', os.path.join(ROOT, 'scripts', 'gen_fortran_api.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, protos = gen.generate()
    assert open(gen.OUT).read() == text, 'run scripts/gen_fortran_api.py'
    assert {name for _, name, _ in protos} == set(L.SYMBOLS)
    for ret, name, args in protos:
        m = re.search(r'function %s\(([^)]*)\)' % name, text)
        assert m, name
        dummies = [a for a in m.group(1).split(',') if a.strip()]
        assert len(dummies) == len(args), name


def _fortran_calls(src):
    """(name, [actual arguments]) of every atx_* function reference in a Fortran source: comments stripped,
    continuation lines joined, arguments split at top-level commas"""
    import re
    lines, cur = [], ''
    for raw in src.splitlines():
        line = raw.split('!')[0].rstrip()
        if not line.strip():
            continue
        body = line.strip()
        if body.startswith('&'):
            body = body[1:]
        if body.endswith('&'):
            cur += body[:-1]
            continue
        lines.append(cur + body)
        cur = ''
    calls = []
    for line in lines:
        for m in re.finditer(r'\b(atx_[a-z0-9_]+)\s*\(', line):
            depth, args, start = 1, [], m.end()
            for pos in range(m.end(), len(line)):
                ch = line[pos]
                if ch == '(':
                    depth += 1
                elif ch == ')':
                    depth -= 1
                    if depth == 0:
                        args.append(line[start:pos].strip())
                        break
                elif ch == ',' and depth == 1:
                    args.append(line[start:pos].strip())
                    start = pos + 1
            else:
                raise AssertionError('unbalanced call: ' + line)
            calls.append((m.group(1), [a for a in args if a]))
    return calls


def test_fortran_shims_call_the_generated_interfaces_consistently():
    """No Fortran compiler exists in this image, so the hand-written shims are checked the way a compiler
    would check them against atx_c_api.f90: every atx_* reference names an existing interface, passes as
    many arguments as it has dummies, passes c_loc()/C_NULL_PTR/c_* handles only where the dummy is a
    type(c_ptr) value, and passes nothing but those where it is"""
    import glob
    import re
    api = open(os.path.join(ROOT, 'atomistica_b200', 'fortran', 'atx_c_api.f90')).read()
    iface = {}
    for m in re.finditer(r'function (atx_[a-z0-9_]+)\(([^)]*)\)(.*?)endfunction', api, re.S):
        names = [a.strip() for a in m.group(2).split(',') if a.strip()]
        kinds = {}
        for decl in m.group(3).splitlines():
            d = re.match(r'\s*(.+?)\s*::\s*(\w+)', decl)
            if d and d.group(2) in names:
                kinds[d.group(2)] = d.group(1)
        iface[m.group(1)] = [(n, kinds[n]) for n in names]
    assert len(iface) == len(L.SYMBOLS)
    shims = sorted(glob.glob(os.path.join(ROOT, 'atomistica_b200', 'fortran', '*_gpu.f90')))
    assert len(shims) == 5
    local = {'atx_pass_error', 'atx_ctx', 'atx_nl', 'atx_p', 'atx_pot'}   # defined by the shims / the module
    assert re.search(r'type\(c_ptr\), save :: atx_ctx = C_NULL_PTR', api)
    ncalls = 0
    for path in shims:
        for name, args in _fortran_calls(open(path).read()):
            if name in local:
                continue
            assert name in iface, '%s: %s has no interface' % (os.path.basename(path), name)
            dummies = iface[name]
            assert len(args) == len(dummies), '%s: %s passes %d arguments, interface has %d' % (
                os.path.basename(path), name, len(args), len(dummies))
            for actual, (dummy, kind) in zip(args, dummies):
                is_ptr_actual = bool(re.match(r'(c_loc\(|C_NULL_PTR|c_[a-z]+\b|atx_ctx\b|[\w%]*atx_(nl|p|pot)\b)', actual))
                if kind.replace(' ', '') == 'type(c_ptr),value':
                    assert is_ptr_actual, '%s: %s(%s=%s): a c_ptr value is expected' % (
                        os.path.basename(path), name, dummy, actual)
                elif kind.strip() == 'type(c_ptr)':        # handle returned by reference
                    assert re.search(r'atx_(ctx|nl|p|pot)\b', actual), (name, dummy, actual)
                else:
                    assert not actual.startswith(('c_loc(', 'C_NULL_PTR')), '%s: %s(%s=%s): interface takes %s' % (
                        os.path.basename(path), name, dummy, actual, kind)
            ncalls += 1
    assert ncalls >= 20


def _fortran_types(api):
    """{type name: [(component, fortran type, dimension)]} of the bind(C) derived types of atx_c_api.f90"""
    import re
    out = {}
    for m in re.finditer(r'type, bind\(C\) :: (\w+)\n(.*?)endtype', api, re.S):
        comps = []
        for line in m.group(2).strip().splitlines():
            d = re.match(r'\s*(\S+) :: (\w+)(?:\((\d+)\))?', line)
            comps.append((d.group(2), d.group(1), int(d.group(3) or 0)))
        out[m.group(1)] = comps
    return out


def test_fortran_derived_types_mirror_the_structs():
    """the bind(C) derived types generated from the header have the components of the ctypes mirrors (which
    test_header_is_plain_c_and_struct_layouts_match_ctypes pins to the compiler's layout): same names, same
    order, same basic type, same array bound"""
    api = open(os.path.join(ROOT, 'atomistica_b200', 'fortran', 'atx_c_api.f90')).read()
    types = _fortran_types(api)
    mirrors = dict(atx_spline_t=L.AtxSpline, atx_bop_params_t=L.AtxBopParams, atx_bop_screening_t=L.AtxBopScreening,
                   atx_juslin_params_t=L.AtxJuslinParams, atx_juslin_screening_t=L.AtxJuslinScreening,
                   atx_pair_params_t=L.AtxPairParams, atx_rebo2_params_t=L.AtxRebo2Params,
                   atx_rebo2_screening_t=L.AtxRebo2Screening)
    assert set(types) == set(mirrors)
    for name, cls in mirrors.items():
        comps = types[name]
        assert len(comps) == len(cls._fields_), name
        for (fname, ftype, dim), (cname, ctype) in zip(comps, cls._fields_):
            assert fname == cname.rstrip('_'), (name, fname, cname)      # ctypes: lambda_ (Python keyword)
            length = getattr(ctype, '_length_', 0)
            base = ctype._type_ if length else ctype
            assert dim == length, (name, fname)
            want = {C.c_int: 'integer(c_int)', C.c_double: 'real(c_double)'}.get(base, 'type(c_ptr)')
            assert ftype == want, (name, fname, ftype, want)
    # Fortran names are case-insensitive: no two components of a type may differ in case only
    for name, comps in types.items():
        lowered = [c[0].lower() for c in comps]
        assert len(set(lowered)) == len(lowered), name
    # the constants the shims select kinds with
    for const, value in (('ATX_BOP_TERSOFF', 1), ('ATX_BOP_KUMAGAI', 2), ('ATX_BOP_BRENNER', 3), ('ATX_BOP_MAX_PAIRS', 6)):
        assert 'parameter :: %s = %d\n' % (const, value) in api


def test_fortran_shims_use_existing_types_and_components():
    """every type(atx_*_t) a shim declares exists in atx_c_api.f90, and every component it assigns or reads
    through such a variable is a component of that type"""
    import glob
    import re
    api = open(os.path.join(ROOT, 'atomistica_b200', 'fortran', 'atx_c_api.f90')).read()
    types = _fortran_types(api)
    nrefs = 0
    for path in sorted(glob.glob(os.path.join(ROOT, 'atomistica_b200', 'fortran', '*_gpu.f90'))):
        src = '\n'.join(line.split('!')[0] for line in open(path).read().splitlines())
        variables = {}
        for m in re.finditer(r'type\((atx_\w+_t)\)[^:\n]*::\s*([^\n]+)', src):
            assert m.group(1) in types, '%s: unknown type %s' % (os.path.basename(path), m.group(1))
            for var in re.findall(r'(\w+)(?:\([^)]*\))?\s*(?:,|$)', m.group(2)):
                variables[var] = m.group(1)
        for var, tname in variables.items():
            comps = {c[0].lower() for c in types[tname]}
            for ref in re.findall(r'\b%s(?:\([^)%%]*\))?%%(\w+)' % var, src):
                assert ref.lower() in comps, '%s: %s%%%s is not a component of %s' % (
                    os.path.basename(path), var, ref, tname)
                nrefs += 1
    assert nrefs >= 80
