#!/usr/bin/env python3
"""Generate atomistica_b200/fortran/atx_c_api.f90 -- one ISO_C_BINDING interface per entry point of
include/atomistica_b200.h -- from the header itself, so the Fortran side cannot fall behind the C ABI.

    python scripts/gen_fortran_api.py            # rewrites the .f90
    python scripts/gen_fortran_api.py --check    # exit 1 when the committed file is stale

Opaque handles and parameter structs travel as type(c_ptr) (c_loc of a bind(C) derived type; the types are
generated from the typedefs of the header, the ATX_* constants from its #defines).  tests/test_abi.py checks
that every exported symbol has an interface with the right number of dummy arguments, that the derived types
equal the ctypes mirrors component by component, and that the hand-written shims use both consistently.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'atomistica_b200.h')
OUT = os.path.join(ROOT, 'atomistica_b200', 'fortran', 'atx_c_api.f90')

SCALAR = {'int': 'integer(c_int)', 'double': 'real(c_double)', 'long long': 'integer(c_long_long)',
          'size_t': 'integer(c_size_t)', 'intptr_t': 'integer(c_intptr_t)', 'char': 'character(kind=c_char)'}
KINDS = {'integer(c_int)': 'c_int', 'real(c_double)': 'c_double', 'integer(c_long_long)': 'c_long_long',
         'integer(c_size_t)': 'c_size_t', 'integer(c_intptr_t)': 'c_intptr_t', 'character(kind=c_char)': 'c_char',
         'type(c_ptr)': 'c_ptr'}


def prototypes(text):
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    text = re.sub(r'#[^\n]*', ' ', text)
    text = re.sub(r'typedef\s+struct\s*\w*\s*\{.*?\}\s*\w+\s*;', ' ', text, flags=re.S)
    text = re.sub(r'typedef[^;]*;', ' ', text)
    text = text.replace('extern "C" {', ' ').replace('}', ' ')
    out = []
    for m in re.finditer(r'([\w\s\*]+?)\b(atx_\w+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), ' '.join(m.group(3).split())
        out.append((ret, name, [] if args in ('', 'void') else [a.strip() for a in args.split(',')]))
    return out


# Pointers to scalars / arrays of a basic type are classified by parameter NAME (the header uses one name for
# one meaning throughout):
#   NULLABLE   optional arguments the header documents as "NULL when absent": type(c_ptr), value -- the caller
#              passes c_loc(x) or C_NULL_PTR (the way the shims in this directory handle Fortran `optional`s);
#   BY_REF     a single scalar result: scalar dummy by reference;
#   the rest   arrays: assumed-size dummy, any rank is passed by sequence association.
NULLABLE = {'mask', 'epot_per_at', 'epot_per_bond', 'f_per_bond', 'wpot_per_at', 'wpot_per_bond', 'already',
            'verlet_shell', 'dvdx', 'dvdy', 'dvdz', 'ms8', 'p2p', 'last_run_ms', 'nrebuilds'}
BY_REF = {'epot', 'ekin', 'range', 'gbs', 'tflops', 'total_ms', 'count', 'nebmax', 'npairs', 'nown', 'nghost',
          'nbuilds', 'nreused'}


def fortran_arg(decl):
    """C parameter declaration -> (name, fortran type, attribute, dimension)"""
    d = decl.replace('const ', '').replace('const*', '*').strip()
    stars = d.count('*')
    d = d.replace('*', ' ')
    toks = d.split()
    name = toks[-1]
    base = ' '.join(toks[:-1])
    if name == 'len':
        name = 'len_'
    if base in SCALAR:
        if stars == 0:
            return name, SCALAR[base], ', value', ''
        if stars == 1:
            if name in NULLABLE:
                return name, 'type(c_ptr)', ', value', ''
            if name in BY_REF:
                return name, SCALAR[base], '', ''
            return name, SCALAR[base], '', '(*)'
        return name, 'type(c_ptr)', '', '(*)'                 # e.g. const int *const *firstneigh
    # void, opaque handles (atx_ctx ...) and parameter structs
    if stars >= 2:
        return name, 'type(c_ptr)', '', ''                     # handle returned through the argument
    return name, 'type(c_ptr)', ', value', ''


def structs(text):
    """typedef struct { ... } name;  ->  [(name, [(field, fortran type, dimension or 0)])], and the integer
    #defines of the header (array bounds, kind selectors)"""
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    defines = []
    for m in re.finditer(r'#define\s+(ATX_\w+)\s+\(?(-?\d+)\)?\s*$', text, flags=re.M):
        defines.append((m.group(1), int(m.group(2))))
    values = dict(defines)
    out = []
    for m in re.finditer(r'typedef\s+struct\s*\{(.*?)\}\s*(atx_\w+)\s*;', text, flags=re.S):
        fields = []
        for stmt in m.group(1).split(';'):
            stmt = ' '.join(stmt.replace('const ', '').split())
            if not stmt:
                continue
            base, rest = stmt.split(' ', 1)
            assert base in ('int', 'double'), stmt
            for item in rest.split(','):
                item = item.strip()
                if item.startswith('*'):
                    fields.append((item.lstrip('* '), 'type(c_ptr)', 0))
                    continue
                a = re.match(r'(\w+)\[(\w+)\]$', item)
                if a:
                    dim = a.group(2)
                    fields.append((a.group(1), SCALAR[base], int(dim) if dim.isdigit() else values[dim]))
                else:
                    assert re.match(r'\w+$', item), stmt
                    fields.append((item, SCALAR[base], 0))
        out.append((m.group(2), fields))
    return out, defines


def generate():
    header = open(HEADER).read()
    protos = prototypes(header)
    types, defines = structs(header)
    lines = ['!! ISO_C_BINDING interface blocks for libatomistica_b200.so -- GENERATED by scripts/gen_fortran_api.py',
             '!! from include/atomistica_b200.h (%d entry points); do not edit.' % len(protos),
             '!! Drop this file into src/support/ of the reference tree; it has no dependencies.  Opaque handles and',
             '!! parameter structs travel as type(c_ptr): pass c_loc() of one of the bind(C) derived types below',
             '!! (atx_bop_params_t, atx_spline_t, ...: the structs of the header, component by component).  No Fortran',
             '!! compiler exists in the build image, so this file and the shims that use it are checked structurally',
             '!! (tests/test_abi.py: against the header, the ctypes mirrors and each other), not compiled.',
             'module atx_c_api', '  use, intrinsic :: iso_c_binding', '  implicit none', '',
             '  !! the device context the shims of one process share (created by the first neighbour-list build)',
             '  type(c_ptr), save :: atx_ctx = C_NULL_PTR', '']
    for name, value in defines:
        lines.append('  integer(c_int), parameter :: %s = %d' % (name, value))
    lines.append('')
    lines.append('  !! bind(C) images of the parameter structs: same components, same order as in the header')
    for name, fields in types:
        lines.append('  type, bind(C) :: %s_t' % name)
        for fname, ftype, dim in fields:
            init = ' = C_NULL_PTR' if ftype == 'type(c_ptr)' else ''
            lines.append('     %s :: %s%s%s' % (ftype, fname, '(%d)' % dim if dim else '', init))
        lines.append('  endtype %s_t' % name)
    lines += ['', '  interface']
    for ret, name, args in protos:
        fa = [fortran_arg(a) for a in args]
        names = [a[0] for a in fa]
        rtype = {'int': 'integer(c_int)', 'long long': 'integer(c_long_long)'}.get(ret, 'type(c_ptr)')
        kinds = sorted({KINDS[a[1]] for a in fa} | {KINDS[rtype]})
        head = '     %s function %s(%s) &' % (rtype, name, ', '.join(names))
        lines.append(head)
        lines.append('          bind(C, name="%s")' % name)
        lines.append('       import :: ' + ', '.join(kinds))
        for n, t, attr, dim in fa:
            lines.append('       %s%s :: %s%s' % (t, attr, n, dim))
        lines.append('     endfunction')
    lines += ['  endinterface', '', 'endmodule atx_c_api', '']
    return '\n'.join(lines), protos


if __name__ == '__main__':
    text, protos = generate()
    if '--check' in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    open(OUT, 'w').write(text)
    print('wrote', OUT, len(protos), 'interfaces')
