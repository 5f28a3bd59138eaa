"""A translator from the Fortran subset the reference's kernels are written in to Python, so that tests can EXECUTE
the reference's own source (no Fortran compiler exists in this image) next to the oracle's restatement of it.  Test
infrastructure only; it reads the sources where they lie under /root/reference and copies nothing.

Subset
  units        [elemental|pure|recursive] subroutine / function; intent(out) dummies become results (a subroutine
               returns the dict of its out / inout arguments), optional dummies default to None, present()
  data         scalars; explicit-shape arrays with arbitrary lower bounds, sections, vector subscripts, whole-array
               and masked (WHERE) assignment, element-wise arithmetic and comparisons (class FA, column-major);
               allocatable components / locals through ALLOCATE / DEALLOCATE / allocated(); derived types as plain
               objects with % components; array constructors; named constants and initialised declarations
  statements   assignment (to an integer variable: truncating conversion), IF blocks and one-line IFs, DO (stride),
               DO WHILE, FORALL (with mask), construct names, CALL with positional / keyword / optional arguments
               of other translated units or of Python callables supplied by the test, RETURN, RAISE_ERROR (an
               exception); INIT_ERROR / PASS_ERROR / write / print are dropped; calls named as no-ops by the test
  cpp          #ifdef / #ifndef / #if defined() with || && ! / #else / #endif; object- and function-like macros
               (macros.inc, filter.inc, spline.inc and the #defines of the kernel files) with token pasting,
               multi-statement bodies (;)
  semantics    IEEE doubles, no contraction; a real literal without kind suffix is default real (single precision,
               promoted); an intent(out) dummy the callee never assigns leaves the actual argument as it was; a
               declared real local starts as NaN so that a use before definition shows up in the results;
               integer / integer in an assignment is refused (it would be an integer division)
  intrinsics   exp sqrt cos sin log acos abs max min int floor real mod-free; dot_product matmul sum (also along a
               dimension) maxval any all shape size lbound ubound spread cross_product iand ishft; PAIR_INDEX macros
Anything outside the subset raises NotImplementedError when the unit is translated (the unit is then recorded as that
exception, and the test that needs it fails) -- nothing is skipped silently."""
import math
import re
import struct


class S:
    """subscript triplet lo:hi (inclusive, None = the bound)"""

    def __init__(self, lo=None, hi=None):
        self.lo, self.hi = lo, hi


class FA:
    """Fortran array: column-major, per-dimension lower bounds (default 1); a(i, j) reads, a[i, j] = x writes;
    sections through S objects"""

    def __init__(self, *shape, data=None, lower=None):
        self.shape = tuple(int(n) for n in shape)
        self.lower = tuple(lower) if lower is not None else (1,) * len(self.shape)
        size = 1
        for n in self.shape:
            size *= n
        self.data = list(data) if data is not None else [0.0] * size
        assert len(self.data) == size, (len(self.data), self.shape)

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data)

    def _select(self, idx):
        """(flat offsets in column-major order, shape of the section or None for one element)"""
        if not isinstance(idx, tuple):
            idx = (idx,)
        assert len(idx) == len(self.shape), (idx, self.shape)
        offs, stride, shape = [0], 1, []
        for i, n, lo in zip(idx, self.shape, self.lower):
            if isinstance(i, FA):                      # vector subscript
                assert all(lo <= int(v) <= lo + n - 1 for v in i.data), (i.data, self.shape)
                r = [int(v) - lo for v in i.data]
                shape.append(len(r))
            elif isinstance(i, S):
                a = lo if i.lo is None else i.lo
                b = lo + n - 1 if i.hi is None else i.hi
                assert b < a or (lo <= a and b <= lo + n - 1), (a, b, self.shape, self.lower)
                r = range(a - lo, b - lo + 1)
                shape.append(len(r))
            else:
                assert lo <= i <= lo + n - 1, (i, self.shape, self.lower)
                r = range(i - lo, i - lo + 1)
            offs = [o + k * stride for k in r for o in offs]
            stride *= n
        return offs, (shape if any(isinstance(i, (S, FA)) for i in idx) else None)

    def __call__(self, *idx):
        offs, shape = self._select(idx)
        if shape is None:
            return self.data[offs[0]]
        return FA(*shape, data=[self.data[o] for o in offs])

    def __getitem__(self, idx):
        return self(*idx) if isinstance(idx, tuple) else self(idx)

    def __setitem__(self, idx, v):
        offs, _ = self._select(idx)
        if isinstance(v, FA):
            assert len(v.data) == len(offs), (len(v.data), len(offs))
            for o, x in zip(offs, list(v.data)):
                self.data[o] = x
        else:
            for o in offs:
                self.data[o] = v

    def assign(self, v):
        """whole-array assignment: scalar broadcast or element-wise copy"""
        if isinstance(v, FA):
            assert len(v.data) == len(self.data)
            self.data[:] = list(v.data)
        else:
            self.data[:] = [v] * len(self.data)

    # the element-wise arithmetic the sources use on sections
    def _zip(self, o, f):
        if isinstance(o, FA):
            assert len(o.data) == len(self.data)
            return FA(*self.shape, data=[f(a, b) for a, b in zip(self.data, o.data)])
        return FA(*self.shape, data=[f(a, o) for a in self.data])

    def __add__(self, o): return self._zip(o, lambda a, b: a + b)
    def __sub__(self, o): return self._zip(o, lambda a, b: a - b)
    def __mul__(self, o): return self._zip(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._zip(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._zip(o, lambda a, b: a / b)
    def __neg__(self): return FA(*self.shape, data=[-a for a in self.data])
    def __pow__(self, o): return self._zip(o, lambda a, b: a ** b)
    def __rsub__(self, o): return self._zip(o, lambda a, b: b - a)
    def __rtruediv__(self, o): return self._zip(o, lambda a, b: b / a)
    def __lt__(self, o): return self._zip(o, lambda a, b: a < b)
    def __gt__(self, o): return self._zip(o, lambda a, b: a > b)
    def __le__(self, o): return self._zip(o, lambda a, b: a <= b)
    def __ge__(self, o): return self._zip(o, lambda a, b: a >= b)
    def __eq__(self, o): return self._zip(o, lambda a, b: a == b)
    def __ne__(self, o): return self._zip(o, lambda a, b: a != b)
    __hash__ = None

    def assign_where(self, mask, v):
        """WHERE (mask) self = v"""
        for k, m in enumerate(mask.data):
            if m:
                self.data[k] = v.data[k] if isinstance(v, FA) else v
    def __radd__(self, o): return self._zip(o, lambda a, b: b + a)


def F1(values):
    return FA(len(values), data=values)


class Obj:
    """derived-type instance; components are set by the test or by translated code"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _set(obj, name, value):
    cur = getattr(obj, name, None)
    if isinstance(cur, FA):
        cur.assign(value)
    else:
        setattr(obj, name, value)


def _sp(x):
    """a real literal without kind suffix is default real: single precision, promoted when used"""
    return struct.unpack('f', struct.pack('f', x))[0]


def _pair_index(i, j, maxval):
    # macros.inc:123
    return 1 + min((i - 1) + (j - 1) * maxval, (j - 1) + (i - 1) * maxval) - min((i - 1) * i // 2, (j - 1) * j // 2)


def _elementwise(f):
    return lambda x, *a: FA(*x.shape, data=[f(v) for v in x.data]) if isinstance(x, FA) else f(x)


def _cross_product(a, b):
    # f_linearalgebra.f90 cross_product: (a2 b3 - a3 b2, a3 b1 - a1 b3, a1 b2 - a2 b1)
    a1, a2, a3 = a.data
    b1, b2, b3 = b.data
    return FA(3, data=[a2 * b3 - a3 * b2, a3 * b1 - a1 * b3, a1 * b2 - a2 * b1])


def _dot_product(a, b):
    acc = 0.0
    for x, y in zip(a.data, b.data):
        acc = acc + x * y
    return acc


def _matmul(a, b):
    """matrix (n, m) times vector (m): each component accumulated over the second index in order"""
    n, m = a.shape
    assert len(b.data) == m
    out = []
    for i in range(n):
        acc = 0.0
        for j in range(m):
            acc = acc + a.data[i + n * j] * b.data[j]
        out.append(acc)
    return FA(n, data=out)


def _outer_product(x, y):
    # macros.inc:82: spread(x, dim=2, ncopies=size(y)) * spread(y, dim=1, ncopies=size(x))
    return FA(len(x.data), len(y.data), data=[a * b for b in y.data for a in x.data])


def _spread(v, dim, ncopies):
    """SPREAD of a vector: dim=1 -> (ncopies, n) with v along the second index; dim=2 -> (n, ncopies)"""
    n = len(v.data)
    if dim == 1:
        return FA(ncopies, n, data=[x for x in v.data for _ in range(ncopies)])
    assert dim == 2
    return FA(n, ncopies, data=list(v.data) * ncopies)


def _sum(a, dim=None):
    if dim is None:
        acc = 0.0
        for x in a.data:
            acc = acc + x
        return acc
    assert len(a.shape) == 2 and dim == 2, (a.shape, dim)       # SUM(array, 2): over the second index, in order
    n, m = a.shape
    out = []
    for i in range(n):
        acc = 0.0
        for j in range(m):
            acc = acc + a.data[i + n * j]
        out.append(acc)
    return FA(n, data=out)


_PI = 3.14159265358979323846264338327950288
INTRINSICS = dict(exp=math.exp, sqrt=math.sqrt, cos=math.cos, sin=math.sin, log=math.log, acos=math.acos,
                  abs=abs, max=max, min=min, real=lambda x, kind=None: float(x), int=_elementwise(int), DP=8,
                  any=lambda a: any(a.data), all=lambda a: all(a.data), shape=lambda a: FA(len(a.shape), data=list(a.shape)),
                  cross_product=_cross_product, maxval=lambda a: max(a.data) if len(a.data) else -2147483648,
                  PAIR_INDEX=_pair_index,
                  PAIR_INDEX_NS=lambda i, j, maxval: j + (i - 1) * maxval,                      # macros.inc:139
                  TRIPLET_INDEX_NS=lambda i, j, k, maxval: k + maxval * (j - 1 + maxval * (i - 1)),  # macros.inc:146
                  floor=_elementwise(math.floor), present=lambda x: x is not None, allocated=lambda x: x is not None,
                  lbound=lambda a, d: a.lower[d - 1], ubound=lambda a, d: a.lower[d - 1] + a.shape[d - 1] - 1, size=lambda a, d=None: len(a) if d is None else a.shape[d - 1],
                  dot_product=_dot_product, matmul=_matmul, spread=_spread, outer_product=_outer_product, sum=_sum,
                  iand=lambda a, b: a & b, ishft=lambda a, n: a << n if n >= 0 else a >> -n,
                  PI=_PI, pi=_PI, Obj=Obj, FA=FA, S=S, _set=_set, _sp=_sp, UNDEF=float('nan'),
                  _new_type=lambda name: Obj(), c_loc=lambda x: x, c_associated=lambda x: x is not None, C_NULL_PTR=None,
                  c_int=4, c_double=8, c_long_long=8, merge=lambda a, b, mask: a if mask else b,
                  reshape=lambda a, shape: FA(*[int(n) for n in shape.data], data=list(a.data)),
                  _ac=lambda v: FA(len(v), data=v))

PY_KEYWORDS = {'lambda': 'lambda_', 'for': 'for_', 'in': 'in_', 'is': 'is_', 'del': 'del_', 'pass': 'pass_'}


def _py(name):
    return PY_KEYWORDS.get(name, name)
ERROR_ARGS = ('error', 'ierror')


def _strip_comment(line):
    quote = None
    for k, ch in enumerate(line):
        if quote:
            if ch == quote:
                quote = None
        elif ch in '"\'':
            quote = ch
        elif ch == '!':
            return line[:k]
    return line


def _cpp_condition(raw, defined, values=None):
    d = raw.split()
    if d[0] == '#ifdef':
        return d[1] in defined
    if d[0] == '#ifndef':
        return d[1] not in defined
    cond = raw[len(d[0]):].split('/*')[0]
    cond = re.sub(r'defined\s*\(\s*(\w+)\s*\)', lambda m: str(m.group(1) in defined), cond)
    for name, value in (values or {}).items():            # object-like macros with a numeric value
        cond = re.sub(r'\b%s\b' % name, str(value), cond)
    cond = cond.replace('||', ' or ').replace('&&', ' and ')
    cond = re.sub(r'!(?!=)', ' not ', cond)
    cond = re.sub(r'\b(?!True\b|False\b|and\b|or\b|not\b)[A-Za-z_]\w*', '0', cond)     # an undefined identifier is 0
    if not re.fullmatch(r'[\s()\w=!<>]*', cond):
        raise NotImplementedError(raw)
    return bool(eval(cond, {'__builtins__': {}}, {}))      # constants, comparisons and boolean operators only


MACRO_SKIP = {'outer_product', 'PAIR_INDEX', 'PAIR_INDEX_NS', 'TRIPLET_INDEX_NS'}    # provided as Python intrinsics


def load_macros(text, defined=()):
    """{name: (parameter list or None, body)} of the #define lines of a header that are active for `defined`"""
    macros, active = {}, []
    for raw in text.splitlines():
        if not raw.startswith('#'):
            continue
        d = raw.split()
        if d[0] in ('#ifdef', '#ifndef', '#if'):
            active.append(_cpp_condition(raw, defined))
        elif d[0] == '#else':
            active[-1] = not active[-1]
        elif d[0] == '#endif':
            active.pop()
        elif d[0] == '#define' and all(active):
            m = re.match(r'#define\s+(\w+)(\(([^)]*)\))?\s*(.*)$', raw)
            if m.group(1) not in MACRO_SKIP:
                params = None if m.group(2) is None else [a.strip() for a in m.group(3).split(',') if a.strip()]
                macros[m.group(1)] = (params, m.group(4).strip())
    return macros


def expand_macros(line, macros):
    for _ in range(30):
        changed = False
        for name, (params, body) in macros.items():
            for m in re.finditer(r'(?<!\w)%s\b' % name, line):        # token replacement, also after %% (as cpp does)
                if params is None:
                    line = line[:m.start()] + body + line[m.end():]
                else:
                    k = m.end()
                    while k < len(line) and line[k] == ' ':
                        k += 1
                    if k >= len(line) or line[k] != '(':
                        continue
                    close = _matching(line, k)
                    actuals = [a.strip() for a in _split_top(line[k + 1:close], ',')]
                    assert len(actuals) == len(params), (name, line)
                    sub = body
                    # simultaneous substitution of the parameters (whole words)
                    sub = re.sub(r'\b(%s)\b' % '|'.join(map(re.escape, params)),
                                 lambda r: actuals[params.index(r.group(1))], sub)
                    sub = re.sub(r'\s*##\s*', '', sub.replace('/**/', ''))        # token pasting (both cpp dialects)
                    line = line[:m.start()] + sub + line[close + 1:]
                changed = True
                break
            if changed:
                break
        if not changed:
            return line
    raise NotImplementedError('macro recursion: ' + line)


def preprocess(text, defined=(), macros=None):
    """cpp conditionals, comments, continuation lines, macro expansion (when `macros` is given), ';' -> list of
    statements"""
    out, active, taken = [], [], []
    values = {k: v[1] for k, v in (macros or {}).items() if v[0] is None and re.fullmatch(r'-?\d+', v[1])}
    for raw in text.splitlines():
        if raw.startswith('#'):
            d = raw.split()
            if d[0] in ('#ifdef', '#ifndef', '#if'):
                active.append(_cpp_condition(raw, defined, values))
                taken.append(active[-1])
            elif d[0] == '#elif':
                active[-1] = (not taken[-1]) and _cpp_condition(raw, defined, values)
                taken[-1] = taken[-1] or active[-1]
            elif d[0] == '#else':
                active[-1] = not taken[-1]
                taken[-1] = True
            elif d[0] == '#endif':
                active.pop(); taken.pop()
            elif d[0] in ('#include', '#define', '#undef'):
                pass
            else:
                raise NotImplementedError(raw)
            continue
        if all(active):
            out.append(_strip_comment(raw).rstrip())
    joined, cur = [], ''
    for line in out:
        body = line.strip()
        if not body:
            continue
        if body.startswith('&'):
            body = body[1:].lstrip()
        if body.endswith('&'):
            cur += body[:-1] + ' '
            continue
        joined.append(cur + body)
        cur = ''
    expanded = []
    for line in joined:
        if macros:
            line = expand_macros(line, macros)
        expanded += [part.strip() for part in _split_statements(line) if part.strip()]
    return expanded


def _split_statements(line):
    """';' separates statements (outside strings)"""
    out, quote, start = [], None, 0
    for k, ch in enumerate(line):
        if quote:
            if ch == quote:
                quote = None
        elif ch in '"\'':
            quote = ch
        elif ch == ';':
            out.append(line[start:k])
            start = k + 1
    out.append(line[start:])
    return out


def _matching(s, start):
    depth, quote = 0, None
    for k in range(start, len(s)):
        ch = s[k]
        if quote:
            if ch == quote:
                quote = None
        elif ch in '"\'':
            quote = ch
        elif ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
            if depth == 0:
                return k
    raise NotImplementedError('unbalanced: ' + s)


def _split_top(s, sep):
    out, depth, start, quote = [], 0, 0, None
    for k, ch in enumerate(s):
        if quote:
            if ch == quote:
                quote = None
        elif ch in '"\'':
            quote = ch
        elif ch in '([':
            depth += 1
        elif ch in ')]':
            depth -= 1
        elif ch == sep and depth == 0:
            out.append(s[start:k])
            start = k + 1
    out.append(s[start:])
    return out


def _sections(e):
    """subscript triplets inside any parenthesised list -> S(lo, hi)"""
    out, k = '', 0
    while k < len(e):
        if e[k] == '(':
            close = _matching(e, k)
            pieces = []
            for piece in _split_top(e[k + 1:close], ','):
                tri = _split_top(piece, ':')
                if len(tri) == 1:
                    pieces.append(_sections(piece))
                elif len(tri) == 2:
                    lo, hi = (_sections(t).strip() or 'None' for t in tri)
                    pieces.append('S(%s, %s)' % (lo, hi))
                else:
                    raise NotImplementedError('stride: ' + e)
            out += '(' + ','.join(pieces) + ')'
            k = close + 1
        else:
            out += e[k]
            k += 1
    return out


def expr(e):
    e = e.replace('(/', '_ac([').replace('/)', '])')
    # default-real literals first (they carry neither a kind suffix nor a D exponent)
    e = re.sub(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])', r'_sp(\1)', e)
    e = re.sub(r'(\d)_DP\b', r'\1', e)
    e = re.sub(r'(?<![\w.])(\d+)_[A-Za-z]\w*', r'\1', e)               # 0_c_int, 2048_c_int
    e = re.sub(r'(\d+\.?\d*|\.\d+)[dD]([-+]?\d+)', r'\1e\2', e)
    for a, b in (('.and.', ' and '), ('.or.', ' or '), ('.not.', ' not '), ('/=', '!='), ('.true.', 'True'),
                 ('.false.', 'False'), ('.le.', '<='), ('.ge.', '>='), ('.lt.', '<'), ('.gt.', '>'), ('.eq.', '=='),
                 ('.ne.', '!=')):
        e = re.sub(re.escape(a), b, e, flags=re.I)
    e = e.replace('%', '.')
    for k, v in PY_KEYWORDS.items():
        e = re.sub(r'\.%s\b' % k, '.' + v, e)
        e = re.sub(r'(?<![\w.])%s\b' % k, v, e)
    # int/int would be an integer division in Fortran: flag it where both operands open a term (after ( + - = ,)
    if re.search(r'(?:^|[(+\-=,])\s*\d+\s*/\s*\d+(?![\w.])', e):
        raise NotImplementedError('integer division: ' + e)
    return _sections(e).strip()


def _lhs(target, arrays):
    """python statement template (with %s for the value) for a scalar, component, element, section or whole array"""
    target = target.strip()
    m = re.fullmatch(r'([\w%]+)\((.*)\)', target)
    if m:
        idx = expr('(' + m.group(2) + ')')[1:-1]
        return '%s[%s] = %%s' % (expr(m.group(1)), idx)
    if '%' in target:
        obj, comp = target.rsplit('%', 1)
        return "_set(%s, '%s', %%s)" % (expr(obj), PY_KEYWORDS.get(comp, comp))
    if target in arrays:
        return '%s.assign(%%s)' % target
    return '%s = %%s' % expr(target)


def _create(name, dims):
    if dims is None:
        return '%s = None' % name          # allocatable: created by ALLOCATE
    shape, lower = [], []
    for d in _split_top(dims, ','):
        if ':' in d:
            lo, hi = d.split(':')
            lower.append(expr(lo)); shape.append('(%s)-(%s)+1' % (expr(hi), expr(lo)))
        else:
            lower.append('1'); shape.append(expr(d))
    return '%s = FA(%s, lower=(%s,))' % (name, ', '.join(shape), ', '.join(lower))


def _results(outputs):
    return 'return dict(%s)' % ', '.join('%s=%s' % (_py(o), _py(o)) for o in outputs if o not in ERROR_ARGS)


def statements(lines, indent=1, outputs=(), sigs=None, arrays=(), local_dims=None, integers=None):
    """translate a list of executable statements; returns python source lines"""
    py, depth = [], indent
    sigs = sigs or {}

    def emit(s):
        py.append('    ' * depth + s)

    in_where = None
    forall_masks = []
    for stmt in lines:
        stmt = re.sub(r'^\w+\s*:\s*(?=(do|if)\b)', '', stmt, flags=re.I)          # construct names
        stmt = re.sub(r'^(end\s*(?:do|if))\s+\w+$', r'\1', stmt, flags=re.I)
        low = stmt.lower()
        if low == 'implicit none' or low.startswith('use ') or re.match(r'(write|print)\b', low):
            continue
        if re.match(r'(invoke_delayed_error)\s*\(', low):
            continue
        m = re.fullmatch(r'do\s+while\s*\((.*)\)', stmt, re.I)
        if m:
            emit('while %s:' % expr(m.group(1))); depth += 1
            emit('pass')
            continue
        m = re.match(r'forall\s*\(', stmt, re.I)
        if m:       # the bodies in the sources have no cross-iteration dependence: a DO loop (+ IF for the mask)
            close = _matching(stmt, m.end() - 1)
            parts = _split_top(stmt[m.end():close], ',')
            var, rng = parts[0].split('=', 1)
            lo, hi = _split_top(rng, ':')
            emit('for %s in range(%s, (%s)+1):' % (var.strip(), expr(lo), expr(hi))); depth += 1
            emit('pass')
            rest = stmt[close + 1:].strip()
            if len(parts) > 1:
                assert len(parts) == 2, stmt
                emit('if %s:' % expr(parts[1]))
                depth += 1
                emit('pass')
                forall_masks.append(depth)
            if rest:
                py.extend(statements([rest], depth, outputs, sigs, arrays))
                depth -= 2 if len(parts) > 1 else 1
                if len(parts) > 1:
                    forall_masks.pop()
            continue
        if re.fullmatch(r'end\s*forall', low):
            if forall_masks and forall_masks[-1] == depth:
                forall_masks.pop()
                depth -= 1
            depth -= 1
            continue
        m = re.fullmatch(r'where\s*\((.*)\)', stmt, re.I)
        if m:
            emit('_mask = %s' % expr(m.group(1)))
            in_where = True
            continue
        if re.fullmatch(r'end\s*where', low):
            in_where = None
            continue
        if in_where:
            k = re.search(r'(?<![=/<>])=(?!=)', stmt).start()
            target = stmt[:k].strip()
            assert re.fullmatch(r'[\w%]+', target), stmt
            emit('%s.assign_where(_mask, %s)' % (expr(target), expr(stmt[k + 1:])))
            continue
        if '::' in stmt and re.match(r'(type\s*\(|integer|real|logical|character)', low):
            # named constants and initialised scalars: "real(DP), parameter :: sig = 0.5"
            is_arg = re.search(r'\bintent\b', stmt.split('::')[0], re.I)
            tm = re.match(r'type\s*\(\s*(\w+)\s*\)', stmt, re.I)
            if tm and not is_arg and local_dims is not None:
                for ent in _split_top(stmt.split('::', 1)[1], ','):      # a derived-type local: an instance of its own
                    name = re.match(r'\s*(\w+)', ent).group(1)
                    if name in local_dims and local_dims[name] is not None:      # an array of them
                        emit(_create(name, local_dims[name]))
                        emit("%s.data[:] = [_new_type('%s') for _ in %s.data]" % (name, tm.group(1), name))
                    else:
                        emit("%s = _new_type('%s')" % (name, tm.group(1)))
                continue
            for ent in _split_top(stmt.split('::', 1)[1], ','):
                name = re.match(r'\s*(\w+)', ent).group(1)
                if local_dims and name in local_dims:       # local arrays come to life where they are declared
                    emit(_create(name, local_dims[name]))
                elif local_dims is not None and not is_arg and '=' not in ent and '(' not in ent \
                        and re.match(r'real', low):
                    emit('%s = UNDEF' % _py(name))          # an undefined real local: visible if it is ever used
                if '=' in ent:
                    emit('%s = %s' % (_py(name), expr(ent.split('=', 1)[1])))
            continue
        m = re.fullmatch(r'deallocate\s*\(([\w%]+)\)', stmt, re.I)
        if m:
            if '%' in m.group(1):
                obj, comp = m.group(1).rsplit('%', 1)
                emit("setattr(%s, '%s', None)" % (expr(obj), comp))
            else:
                emit('%s = None' % m.group(1))
            continue
        m = re.fullmatch(r'allocate\s*\((.*)\)', stmt, re.I)
        if m:
            for ent in _split_top(m.group(1), ','):
                a = re.fullmatch(r'\s*([\w%]+)\((.*)\)\s*', ent)
                if not a:
                    raise NotImplementedError(stmt)
                target, dims = a.group(1), expr(a.group(2))
                if '%' in target:
                    obj, comp = target.rsplit('%', 1)
                    emit("setattr(%s, '%s', FA(%s))" % (expr(obj), comp, dims))
                else:
                    emit('%s = FA(%s)' % (target, dims))
            continue
        if re.match(r'(init_error|pass_error|pass_error_with_info)\s*\(', low):
            continue          # the error stack: an error is an exception here
        if re.fullmatch(r'end\s*if', low) or re.fullmatch(r'end\s*do', low):
            depth -= 1
            continue
        if low == 'else':
            depth -= 1; emit('else:'); depth += 1
            emit('pass')
            continue
        if low == 'return':
            emit(_results(outputs))
            continue
        m = re.match(r'(else\s*if|if)\s*\(', low)
        if m:
            close = _matching(stmt, m.end() - 1)
            cond, rest = expr(stmt[m.end():close]), stmt[close + 1:].strip()
            kw = 'elif' if low.startswith('else') else 'if'
            if rest.lower() == 'then':
                if kw == 'elif':
                    depth -= 1
                emit('%s %s:' % (kw, cond)); depth += 1
                emit('pass')
            else:
                assert kw == 'if', stmt
                emit('if %s:' % cond)
                depth += 1
                py.extend(statements([rest], depth, outputs, sigs, arrays, None, integers))
                depth -= 1
            continue
        m = re.fullmatch(r'do\s+(\w+)\s*=\s*(.+)', stmt, re.I)
        if m:
            parts = [expr(p) for p in _split_top(m.group(2), ',')]
            if len(parts) == 2:
                emit('for %s in range(%s, (%s)+1):' % (m.group(1), parts[0], parts[1]))
            else:
                step = int(parts[2])
                emit('for %s in range(%s, (%s)%s, %d):' % (m.group(1), parts[0], parts[1], '+1' if step > 0 else '-1', step))
            depth += 1
            emit('pass')
            continue
        m = re.match(r'RAISE_ERROR\s*\((.*)\)$', stmt)     # before assignments: the message may hold '='
        if m:
            emit('raise RuntimeError(%r)' % m.group(1))
            continue
        m = re.fullmatch(r'call\s+(\w+)\s*(?:\((.*)\))?', stmt, re.I)
        if m and sigs.get(m.group(1), 0) is None:
            continue                                  # logging / timers: declared as no-ops by the test
        if m:
            name = m.group(1)
            if name not in sigs:
                raise NotImplementedError('call of an unknown unit: ' + stmt)
            dummies, pure_out, outs = sigs[name]
            pairs = []                                     # (dummy, actual), keyword arguments matched by name
            for pos, a in enumerate(a.strip() for a in _split_top(m.group(2) or '', ',') if a.strip()):
                kw = re.fullmatch(r'(\w+)\s*=(?!=)\s*(.+)', a)
                if kw and kw.group(1) in dummies:
                    pairs.append((kw.group(1), kw.group(2)))
                else:
                    assert pos < len(dummies) and not kw, stmt
                    pairs.append((dummies[pos], a))
            pairs = [(d, a) for d, a in pairs if d not in ERROR_ARGS]
            ins = ['%s=%s' % (PY_KEYWORDS.get(d, d), expr(a)) for d, a in pairs if d not in pure_out]
            emit('_r = %s(%s)' % (name, ', '.join(ins)))
            for d, a in pairs:
                if d in outs:
                    # a dummy the callee never assigned leaves the actual argument as it was
                    emit("if _r['%s'] is not None:" % _py(d))
                    emit('    ' + _lhs(a, arrays) % ("_r['%s']" % _py(d)))
            continue
        if '=' in stmt and not low.startswith(('write', 'print', 'select', 'where', 'forall')):
            # first '=' that is not part of ==, /=, <=, >=
            k = re.search(r'(?<![=/<>])=(?!=)', stmt).start()
            fm = re.fullmatch(r'\s*(\w+)\s*\((.*)\)\s*', stmt[k + 1:])
            if fm and sigs.get(fm.group(1)) and sigs[fm.group(1)][2]:
                # a function with intent(out) arguments (the C ABI seen from Fortran): results come back in a dict
                name = fm.group(1)
                dummies, pure_out, outs = sigs[name]
                actuals = [a.strip() for a in _split_top(fm.group(2), ',')]
                assert len(actuals) == len(dummies), stmt
                emit('_r = %s(%s)' % (name, ', '.join(expr(a) for a, d in zip(actuals, dummies) if d not in pure_out)))
                for a, d in zip(actuals, dummies):
                    if d in outs:
                        emit(_lhs(a, arrays) % ("_r['%s']" % _py(d)))
                emit(_lhs(stmt[:k], arrays) % "_r['result']")
                continue
            value = expr(stmt[k + 1:])
            if integers is not None:
                # integer / integer is an integer division in Fortran and a true division in Python: none may occur
                for a, b in re.findall(r'(?<![\w.%)*])(?<!\*\s)(\w+)\s*/\s*(\w+)(?![\w.(%])', stmt[k + 1:]):
                    if (a in integers or a.isdigit()) and (b in integers or b.isdigit()):
                        raise NotImplementedError('integer division: ' + stmt)
            if integers is not None and re.match(r'\s*(\w+)', stmt[:k]).group(1) in integers:
                value = 'int(%s)' % value                 # assignment to an integer variable converts (truncates)
            emit(_lhs(stmt[:k], arrays) % value)
            continue
        raise NotImplementedError(stmt)
    return py


def _signature(lines, k):
    m = re.match(r'(?:(?:elemental|pure|recursive)\s+)*(subroutine|function)\s+(\w+)\s*\(([^)]*)\)', lines[k], re.I)
    if not m:
        return None
    kind, name, args = m.group(1).lower(), m.group(2), [a.strip() for a in m.group(3).split(',') if a.strip()]
    # the unit ends at its own END; internal procedures (after CONTAINS) nest
    depth, end, contains = 0, None, None
    for j in range(k + 1, len(lines)):
        if re.match(r'(?:(?:elemental|pure|recursive)\s+)*(subroutine|function)\s+\w+\s*\(', lines[j], re.I):
            depth += 1
        elif re.match(r'end\s*(subroutine|function)', lines[j], re.I):
            if depth == 0:
                end = j
                break
            depth -= 1
        elif lines[j].strip().lower() == 'contains' and depth == 0:
            contains = j
    body = lines[k + 1:contains if contains is not None else end]
    outs, pure_out, objects, local_arrays, optional, integers = [], set(), [], [], set(), set()
    for decl in body:
        if '::' not in decl or not re.match(r'(type\s*\(|integer|real|logical|character)', decl.lower()):
            continue
        names = [v.strip() for v in _split_top(decl.split('::', 1)[1], ',')]
        bare = [re.match(r'\w+', n).group(0) for n in names]
        if decl.lower().startswith('integer'):
            integers.update(bare)
        if re.search(r'\boptional\b', decl.split('::')[0], re.I):
            optional.update(bare)
        m2 = re.search(r'intent\s*\(\s*(out|inout)\s*\)', decl, re.I)
        if m2:
            outs += bare
            if m2.group(1).lower() == 'out':
                pure_out.update(bare)
                if decl.lower().startswith('type'):
                    objects += bare                 # a derived-type result starts as an empty object
        for n, b in zip(names, bare):
            dims = re.fullmatch(r'\w+\((.*)\)', n.split('=')[0].strip())
            if dims and (b not in args or b in pure_out):
                deferred = all(d.strip() == ':' for d in _split_top(dims.group(1), ','))
                local_arrays.append((b, None if deferred else dims.group(1)))   # locals / array results are created here
    return dict(kind=kind, name=name, args=args, end=end, body=body, outs=outs, pure_out=pure_out, objects=objects,
                local_arrays=local_arrays, optional=optional, integers=integers, contains=contains)


def units(text, defined=(), env=None, macros=None, global_arrays=(), noops=(), global_scalars=()):
    """{name: python callable} for every subroutine / function of a source text.  A subroutine returns the dict
    of its intent(out) / intent(inout) arguments, a function its result.  env: extra names (constants, Python
    callables; a callable that is CALLed needs .fortran_args = (dummy names, names of the intent(out) ones) and
    returns the dict of those)"""
    lines = preprocess(text, defined, macros)
    scope = dict(INTRINSICS)
    scope.update(env or {})
    found, k = [], 0
    while k < len(lines):
        sig = _signature(lines, k)
        if sig is None:
            k += 1
            continue
        found.append(sig)
        if sig['contains'] is not None:          # internal procedures become units of their own (they use their arguments only)
            j = sig['contains'] + 1
            while j < sig['end']:
                inner = _signature(lines, j)
                if inner is None:
                    j += 1
                    continue
                found.append(inner)
                j = inner['end'] + 1
        k = sig['end'] + 1
    sigs = {}
    for name, fn in scope.items():
        if hasattr(fn, 'fortran_args'):
            d, po = fn.fortran_args
            sigs[name] = (list(d), set(po), list(po))
    for sig in found:
        sigs[sig['name']] = (sig['args'], sig['pure_out'], sig['outs'])
    for name in noops:
        sigs[name] = None
    sources = {}
    for sig in found:
        name = sig['name']
        pyargs = [PY_KEYWORDS.get(a, a) + ('=None' if a in sig['optional'] else '')
                  for a in sig['args'] if a not in sig['pure_out']]                       # intent(out): results only
        src = ['def %s(%s):' % (name, ', '.join(pyargs))] + ['    global %s' % g for g in global_scalars] + \
            ['    %s = Obj()' % o for o in sig['objects']]
        arrays = [b for b, _ in sig['local_arrays']] + list(global_arrays)
        # an intent(out) scalar a branch never assigns is undefined in Fortran: None here
        src += ['    %s = None' % _py(o) for o in sig['pure_out'] if o not in sig['objects'] and o not in arrays]
        locals_only = [b for b, _ in sig['local_arrays']]
        try:
            src += statements(sig['body'], 1, sig['outs'] if sig['kind'] == 'subroutine' else (), sigs, arrays,
                              dict(sig['local_arrays']), sig['integers'])
            src.append('    return %s' % name if sig['kind'] == 'function' else '    ' + _results(sig['outs']))
            compile('\n'.join(src), name, 'exec')
            sources[name] = '\n'.join(src)
        except NotImplementedError as e:          # a unit outside the subset: recorded, never silently replaced
            sources[name] = e
        except Exception as e:                    # noqa: BLE001 -- statements the patterns above misread
            sources[name] = NotImplementedError('%s: %r' % (name, e))
    # one namespace for all units of the file, so that they can call each other
    result = {}
    for name, src in sources.items():
        if isinstance(src, str):
            exec(src, scope)
            fn = scope[name]
            fn.python_source = src
            fn.fortran_args = (sigs[name][0], sigs[name][1])
            result[name] = fn
        else:
            result[name] = src
    return result


def run_fragment(text, first, last, env, defined=(), arrays=(), macros=None):
    """execute the statements of `text` from the first one matching regex `first` to the first one after it
    matching `last` (inclusive) in `env` (a dict: loop variables, `this`, ...)"""
    lines = preprocess(text, defined, macros)
    a = next(k for k, s in enumerate(lines) if re.search(first, s))
    b = next(k for k in range(a, len(lines)) if re.search(last, lines[k]))
    sigs = {}
    for name, fn in env.items():
        if hasattr(fn, 'fortran_args'):
            d, po = fn.fortran_args
            sigs[name] = (list(d), set(po), list(po))
    src = statements(lines[a:b + 1], 0, (), sigs, arrays)
    scope = dict(INTRINSICS)
    scope.update(env)
    exec('\n'.join(src), scope)
    return '\n'.join(src)
