"""Start geometries for the hydrocarbons that the reference's tests/test_rebo2_molecules.py takes from
ASE's G2 collection (ase.build.molecule), which is not available here: a carbon skeleton plus
hydrogens at the ideal sp3 / sp2 / sp directions.  Every geometry is relaxed before its energy is
compared with Brenner's table, so only the topology matters."""
import numpy as np

def _unit(v):
    v = np.asarray(v, dtype=float)
    return v / np.linalg.norm(v)

def _perp(u):
    a = np.array([1.0, 0, 0]) if abs(u[0]) < 0.9 else np.array([0, 1.0, 0])
    return _unit(np.cross(u, a))

def hydrogenate(carbons, bonds, hyb, rch=1.09):
    """carbon skeleton (positions, bonds as index pairs, hybridisation 'sp3'/'sp2'/'sp' per atom) ->
    symbols, positions with hydrogens at the ideal directions"""
    C = np.array(carbons, dtype=float)
    nb = {i: [] for i in range(len(C))}
    for a, b in bonds:
        nb[a].append(b); nb[b].append(a)
    H = []
    for i, h in enumerate(hyb):
        u = [_unit(C[j] - C[i]) for j in nb[i]]
        want = dict(sp3=4, sp2=3, sp=2)[h] - len(u)
        if want <= 0:
            continue
        if h == 'sp':
            d = [-u[0]]
        elif h == 'sp2':
            if len(u) == 2:
                d = [-_unit(u[0] + u[1])]
            else:
                # plane from the neighbour's other bonds, else arbitrary
                j = nb[i][0]
                others = [_unit(C[k] - C[j]) for k in nb[j] if k != i]
                n = _unit(np.cross(u[0], others[0])) if others and np.linalg.norm(np.cross(u[0], others[0])) > 1e-6 \
                    else _perp(u[0])
                t = np.cross(n, u[0])
                d = [-0.5 * u[0] + s * np.sqrt(3) / 2 * t for s in (1, -1)]
        else:
            if len(u) == 3:
                d = [-_unit(u[0] + u[1] + u[2])]
            elif len(u) == 2:
                b = -_unit(u[0] + u[1]); n = _unit(np.cross(u[0], u[1]))
                th = np.radians(109.47 / 2)
                d = [np.cos(th) * b + s * np.sin(th) * n for s in (1, -1)]
            else:
                p = _perp(u[0]); q = np.cross(u[0], p)
                c, s_ = np.cos(np.radians(109.47)), np.sin(np.radians(109.47))
                d = [c * u[0] + s_ * (np.cos(a) * p + np.sin(a) * q) for a in (0, 2 * np.pi / 3, 4 * np.pi / 3)]
                if want == 4:      # methane-like: not used
                    d.append(-u[0])
        for v in d[:want]:
            H.append(C[i] + rch * _unit(v))
    return ['C'] * len(C) + ['H'] * len(H), [list(map(float, x)) for x in np.vstack([C, np.array(H)])]

def zigzag(n, r=1.53, ang=111.0):
    a = np.radians(ang / 2)
    return [[k * r * np.sin(a), (k % 2) * r * np.cos(a), 0.0] for k in range(n)]

def extra_molecules():
    m = {}
    m['CH2_s1A1d'] = (['C', 'H', 'H'], [[0, 0, 0], [0.863, 0.699, 0], [-0.863, 0.699, 0]])
    m['C3H8'] = hydrogenate(zigzag(3), [(0, 1), (1, 2)], ['sp3'] * 3)
    m['trans-butane'] = hydrogenate(zigzag(4), [(0, 1), (1, 2), (2, 3)], ['sp3'] * 4)
    t = 1.53 / np.sqrt(3)
    m['isobutane'] = hydrogenate([[0, 0, 0], [t, t, t], [-t, -t, t], [-t, t, -t]], [(0, 1), (0, 2), (0, 3)],
                                 ['sp3'] * 4)
    m['C3H6_Cs'] = hydrogenate([[0, 0, 0], [1.16, 0.67, 0], [2.46, -0.08, 0]], [(0, 1), (1, 2)],
                               ['sp2', 'sp2', 'sp3'])
    m['C3H6_D3h'] = hydrogenate([[0.87, 0, 0], [-0.435, 0.753, 0], [-0.435, -0.753, 0]], [(0, 1), (1, 2), (0, 2)],
                                ['sp3'] * 3)
    m['C3H4_C2v'] = hydrogenate([[0.65, 0, 0], [-0.65, 0, 0], [0, 1.36, 0]], [(0, 1), (1, 2), (0, 2)],
                                ['sp2', 'sp2', 'sp3'])
    m['butadiene'] = hydrogenate([[0, 0, 0], [1.16, 0.67, 0], [2.43, -0.06, 0], [3.59, 0.61, 0]],
                                 [(0, 1), (1, 2), (2, 3)], ['sp2'] * 4)
    m['2-butyne'] = hydrogenate([[0, 0, 0], [1.46, 0, 0], [2.67, 0, 0], [4.13, 0, 0]], [(0, 1), (1, 2), (2, 3)],
                                ['sp3', 'sp', 'sp', 'sp3'])
    return m
