#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's own test DATA files.

Run once in the build container (the only place /root/reference exists):

    python tests/golden/make_fixtures.py

Nothing in tests/, bench.py or smoke() reads /root/reference at run time; they
read the .npz/.json files written here.  Only data (tables, coordinates and the
known-answer numbers the reference's tests assert) is converted -- no reference
source code is copied.

Sources (all under /root/reference/tests):
  Cu_mishin1.eam.alloy, Au-Grochola-JCP05.eam.alloy  -> setfl tables (float64 arrays)
  aC.cfg, aC_small.cfg                               -> cell + cartesian positions
  eam_crash1.poscar, eam_crash2.poscar               -> cell + positions
  molecule_database/*.xyz                            -> molecules.json
  test_rebo2_molecules.py:52-93, test_pbc.py:57-59,
  test_bulk_properties.py:55-167                     -> kat.json (known answers)
"""
import json
import os
import re
import sys

import numpy as np

REF = '/root/reference/tests'
OUT = os.path.dirname(os.path.abspath(__file__))


def read_setfl(fn):
    """setfl layout as parsed by tabulated_alloy_eam.f90:147-259."""
    with open(fn) as f:
        comments = [f.readline().rstrip('\n') for _ in range(3)]
        toks = f.readline().split()
        nel = int(toks[0])
        names = toks[1:1 + nel]
        toks = f.readline().split()
        nF, dF, nr, dr, cutoff = int(toks[0]), float(toks[1]), int(toks[2]), float(toks[3]), float(toks[4])
        rest = f.read().split()
    pos = 0
    Z, mass, a0, lattice, F, rho = [], [], [], [], [], []
    for _ in range(nel):
        Z.append(int(rest[pos])); mass.append(float(rest[pos + 1]))
        a0.append(float(rest[pos + 2])); lattice.append(rest[pos + 3])
        pos += 4
        F.append(np.array(rest[pos:pos + nF], dtype=np.float64)); pos += nF
        rho.append(np.array(rest[pos:pos + nr], dtype=np.float64)); pos += nr
    rphi = []
    for i in range(nel):
        for j in range(i + 1):
            rphi.append(np.array(rest[pos:pos + nr], dtype=np.float64)); pos += nr
    assert pos == len(rest), (pos, len(rest))
    return dict(comments=np.array(comments), names=np.array(names), nF=nF, dF=dF, nr=nr, dr=dr,
                cutoff=cutoff, Z=np.array(Z), mass=np.array(mass), a0=np.array(a0),
                lattice=np.array(lattice), F=np.array(F), rho=np.array(rho), rphi=np.array(rphi))


def read_cfg(fn):
    """AtomEye extended cfg (as read by atomistica.io / ase): H0 rows are the cell vectors."""
    H0 = np.zeros((3, 3))
    s, sym = [], []
    with open(fn) as f:
        lines = f.read().split('\n')
    n = int(lines[0].split('=')[1])
    cur_sym = None
    for l in lines[1:]:
        m = re.match(r'H0\((\d),(\d)\)\s*=\s*([-\d.eE+]+)', l)
        if m:
            H0[int(m.group(1)) - 1, int(m.group(2)) - 1] = float(m.group(3))
            continue
        t = l.split()
        if len(t) == 1 and re.match(r'^[A-Z][a-z]?$', t[0]):
            cur_sym = t[0]
        elif len(t) >= 6 and cur_sym is not None:
            s.append([float(t[0]), float(t[1]), float(t[2])])
            sym.append(cur_sym)
    s = np.array(s)
    assert len(s) == n, (len(s), n)
    # cartesian = s . H0  (rows of H0 are the cell vectors)
    return dict(cell=H0, scaled=s, positions=s @ H0, symbols=np.array(sym))


def read_poscar(fn):
    with open(fn) as f:
        lines = f.read().split('\n')
    sym = lines[0].split()
    scale = float(lines[1])
    cell = np.array([[float(x) for x in lines[2 + i].split()] for i in range(3)]) * scale
    counts = [int(x) for x in lines[5].split()]
    assert lines[6].strip().lower().startswith('c')
    n = sum(counts)
    pos = np.array([[float(x) for x in lines[7 + i].split()[:3]] for i in range(n)]) * scale
    symbols = sum([[s] * c for s, c in zip(sym, counts)], [])
    return dict(cell=cell, positions=pos, symbols=np.array(symbols))


def read_xyz(fn):
    with open(fn) as f:
        lines = f.read().split('\n')
    n = int(lines[0])
    sym, pos = [], []
    for l in lines[2:2 + n]:
        t = l.split()
        sym.append(t[0]); pos.append([float(x) for x in t[1:4]])
    return sym, pos


def read_funcfl(fn):
    """funcfl layout as parsed by tabulated_eam_init (tabulated_eam.f90:172-197)"""
    with open(fn) as f:
        comment = f.readline().rstrip('\n')
        t = f.readline().split()
        Z, mass, a0, lattice = int(t[0]), float(t[1]), float(t[2]), t[3]
        t = f.readline().replace('D', 'E').split()
        nF, dF, nr, dr, cutoff = int(t[0]), float(t[1]), int(t[2]), float(t[3]), float(t[4])
        rest = np.array(f.read().replace('D', 'E').split(), dtype=np.float64)
    return dict(comment=comment, name='Au' if Z == 79 else str(Z), Znum=Z, mass=mass, a0=a0, lattice=lattice,
                nF=nF, dF=dF, nr=nr, dr=dr, cutoff=cutoff, F=rest[:nF], Z=rest[nF:nF + nr],
                rho=rest[nF + nr:nF + 2 * nr])


def main():
    if not os.path.isdir(REF):
        sys.exit('reference tests directory not present; fixtures are already committed')
    np.savez_compressed(os.path.join(OUT, 'au_u3_funcfl.npz'), **read_funcfl(os.path.join(REF, 'Au_u3.eam')))
    for src, dst in [('Cu_mishin1.eam.alloy', 'cu_mishin1_setfl.npz'),
                     ('Au-Grochola-JCP05.eam.alloy', 'au_grochola_setfl.npz')]:
        np.savez_compressed(os.path.join(OUT, dst), **read_setfl(os.path.join(REF, src)))
    for src, dst in [('aC.cfg', 'aC.npz'), ('aC_small.cfg', 'aC_small.npz')]:
        np.savez_compressed(os.path.join(OUT, dst), **read_cfg(os.path.join(REF, src)))
    for src, dst in [('eam_crash1.poscar', 'eam_crash1.npz'), ('eam_crash2.poscar', 'eam_crash2.npz')]:
        np.savez_compressed(os.path.join(OUT, dst), **read_poscar(os.path.join(REF, src)))
    mols = {}
    for fn in sorted(os.listdir(os.path.join(REF, 'molecule_database'))):
        if fn.endswith('.xyz'):
            sym, pos = read_xyz(os.path.join(REF, 'molecule_database', fn))
            mols[fn[:-4]] = dict(symbols=sym, positions=pos)
    json.dump(mols, open(os.path.join(OUT, 'molecules.json'), 'w'), indent=0)

    kat = {
        # tests/test_rebo2_molecules.py:52-93 (Brenner et al. 2002, Table 12), tolerance 0.005 eV
        'rebo2_atomization_eV': {
            'CH2_s1A1d': -8.4693, 'CH3': -13.3750, 'CH4': -18.1851, 'C2H': -11.5722,
            'C2H2': -17.5651, 'C2H4': -24.4077, 'H3C2H2': -26.5601, 'C2H6': -30.8457,
            'C3H4_C2v': -28.2589, 'CH2=C=CH2': -30.2392, 'propyne': -30.3076,
            'C3H6_D3h': -36.8887, 'C3H6_Cs': -37.3047, 'C3H8': -43.5891,
            'butadiene': -43.0035, 'CH3CH=C=CH2': -43.1367, '1-butyne': -43.0510,
            '2-butyne': -43.0501, '1-butene': -50.0487, 'cis-butene': -50.2017,
            'i-C4H9': -52.0451, 't-C4H9': -52.3778, 'trans-butane': -56.3326,
            'isobutane': -56.3309, '1,3-pentadiene': -55.9025, '1,4-pentadiene': -56.5078,
            'cyclopentene': -57.1119, 'cyclopentane': -63.6443, '2-pentene': -62.9456,
            '1-butene,2-methyl': -62.9658, 'n-pentane': -69.0761, 'isopentane': -69.0739,
            'neopentane': -69.0614, 'C6H6': -59.3096, 'cyclohexane': -76.4606,
            'naphthalene': -93.8784},
        'rebo2_atomization_tol_eV': 0.005,
        # tests/test_pbc.py:42-59
        'tersoff_si100_surface_energy_J_m2': 2.309,
        'tersoff_si100_surface_energy_tol': 0.001,
        # tests/test_bulk_properties.py (5 % tolerance): Ec eV, a0 A, C11 C12 C44 GPa
        'bulk': {
            'Tersoff_dia_Si': dict(Ec=4.63, a0=5.432, C11=143.0, C12=75.0, C44=69.0, C440=119.0, B=98.0),
            'Tersoff_dia_C': dict(Ec=7.396 - 0.0250, a0=3.566, C11=1067.0, C12=104.0, C44=636.0, C440=671.0),
            'Tersoff_B3_SiC': dict(Ec=6.165, a0=4.321, C11=437.0, C12=118.0, C440=311.0, B=224.0),
            'Kumagai_dia_Si': dict(Ec=4.630, a0=5.429, C11=166.4, C12=65.3, C440=120.9),
            'Brenner_Erhart_dia_C': dict(Ec=7.3731, a0=3.566, C11=1082.0, C12=127.0, C44=673.0, B=445.0),
            'Brenner_Erhart_dia_Si': dict(Ec=4.63, a0=5.429, C11=167.0, C12=65.0, C440=105.0, B=99.0),
            'Brenner_Erhart_B3_SiC': dict(Ec=6.340, a0=4.359, C11=382.0, C12=145.0, C440=305.0, B=224.0),
            'Rebo2_dia_C': dict(Ec=7.370, a0=3.566, C11=1080.0, C12=130.0, C44=720.0),
            'TabulatedAlloyEAM_fcc_Au': dict(Ec=3.924, a0=4.070, C11=202.0, C12=170.0, C44=47.0, C440=46.0),
            'TabulatedEAM_fcc_Au': dict(Ec=3.93, a0=4.08, B=167.0, C11=183.0, C12=159.0, C44=45.0),
            # Juslin rows, tests/test_bulk_properties.py:95-120 (sc-W there uses a doubled cell: a0 = 2*2.671)
            'Juslin_bcc_W': dict(Ec=8.89, a0=3.165, C11=542.0, C12=191.0, C44=162.0, B=308.0),
            'Juslin_fcc_W': dict(Ec=8.89 - 0.346, a0=4.005),
            'Juslin_sc_W': dict(Ec=8.89 - 1.614, a0=2.671),
            'Juslin_dia_C': dict(Ec=7.376 - 0.0524, a0=3.558, C11=621.0, C12=415.0, C44=383.0, B=484.0),
            'Juslin_B1_WC': dict(Ec=(16.68 - 0.98) / 2, a0=4.380, B=433.0),
            'Juslin_B2_WC': dict(Ec=(16.68 - 2.32) / 2, a0=2.704, B=411.0),
            'Juslin_B3_WC': dict(Ec=(16.68 - 2.12) / 2, a0=4.679, B=511.0),
            # tests/test_bulk_properties.py:72-76 (Brenner II) and :163-177 (Matsunaga B-C-N, also TersoffScr)
            'Brenner_II_dia_C': dict(Ec=7.376 - 0.0524, a0=3.558, C11=621.0, C12=415.0, C44=383.0, B=484.0),
            'Tersoff_BCN_dia_C': dict(Ec=7.396 - 0.0250, a0=3.566, C11=1067.0, C12=104.0, C44=636.0, C440=671.0),
            'Tersoff_BCN_B3_BN': dict(Ec=6.63, a0=3.658, B=385.0),
        },
        'bulk_tol_rel': 0.05,
        'surface_tol_rel': 0.05,
        # all rows of tests/test_surface_properties.py:213-246 that carry a value (r_Jm2): ideal (111),
        # (110), (100) terminations and the dimerised (100)-2x1, relaxed to fmax = 0.005 eV/A
        'surface_relaxed_J_m2': {
            'Brenner': {'C': {'111': 2.06, '110': 2.96, '100': 5.59, '100-2x1': 5.65},
                        'Si': {'111': 0.999, '110': 1.23, '100': 1.95, '100-2x1': 1.13},
                        'SiC': {'111': 1.67, '110': 2.29, '100': 3.93, '100-2x1': 2.85}},
            'BrennerScr': {'C': {'111': 2.06, '110': 2.96, '100': 5.88, '100-2x1': 5.89},
                           'Si': {'111': 0.999, '110': 1.23, '100': 1.90, '100-2x1': 1.13},
                           'SiC': {'111': 1.67, '110': 2.29, '100': 3.87, '100-2x1': 2.91}}},
        # closed form from the reference formulas
        'tersoff_si_diamond_a0_5.432_eV_per_atom': -4.6295950127,
        'kumagai_si_diamond_a0_5.429_eV_per_atom': -4.6299992839,
    }
    json.dump(kat, open(os.path.join(OUT, 'kat.json'), 'w'), indent=1)
    print('fixtures written to', OUT)


if __name__ == '__main__':
    main()
