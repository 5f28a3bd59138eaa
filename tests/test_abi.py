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
