"""File readers/writers needed by the hot path (setfl tables for TabulatedAlloyEAM)."""
import numpy as np


def read_setfl(fn):
    """DYNAMO setfl layout as parsed by tabulated_alloy_eam_init
    (src/potentials/eam/tabulated_alloy_eam.f90:147-259): 3 comment lines; `nel names`;
    `nF dF nr dr cutoff`; per element `Z mass a0 lattice`, nF F values, nr rho values; then the
    lower-triangle r*phi(r) tables."""
    with open(fn) as f:
        comments = [f.readline().rstrip('\n') for _ in range(3)]
        toks = f.readline().split()
        nel = int(toks[0])
        names = toks[1:1 + nel]
        toks = f.readline().replace('D', 'E').split()
        nF, dF, nr, dr, cutoff = int(toks[0]), float(toks[1]), int(toks[2]), float(toks[3]), float(toks[4])
        rest = f.read().split()
    pos = 0
    Z, mass, a0, lattice, F, rho = [], [], [], [], [], []
    for _ in range(nel):
        Z.append(int(rest[pos])); mass.append(float(rest[pos + 1])); a0.append(float(rest[pos + 2]))
        lattice.append(rest[pos + 3])
        pos += 4
        F.append(np.array(rest[pos:pos + nF], dtype=np.float64)); pos += nF
        rho.append(np.array(rest[pos:pos + nr], dtype=np.float64)); pos += nr
    rphi = []
    for i in range(nel):
        for j in range(i + 1):
            rphi.append(np.array(rest[pos:pos + nr], dtype=np.float64)); pos += nr
    return dict(comments=comments, names=names, nF=nF, dF=dF, nr=nr, dr=dr, cutoff=cutoff, Z=np.array(Z),
                mass=np.array(mass), a0=np.array(a0), lattice=lattice, F=np.array(F), rho=np.array(rho),
                rphi=np.array(rphi))


def write_setfl(fn, tab):
    """Inverse of read_setfl with round-trip exact ('%.17g') numbers."""
    with open(fn, 'w') as f:
        for c in list(tab['comments'])[:3]:
            f.write(str(c) + '\n')
        names = [str(x) for x in tab['names']]
        f.write('%d %s\n' % (len(names), ' '.join(names)))
        f.write('%d %.17g %d %.17g %.17g\n' % (int(tab['nF']), float(tab['dF']), int(tab['nr']), float(tab['dr']),
                                               float(tab['cutoff'])))
        for i in range(len(names)):
            f.write('%d %.17g %.17g %s\n' % (int(tab['Z'][i]), float(tab['mass'][i]), float(tab['a0'][i]),
                                             str(tab['lattice'][i])))
            for arr in (tab['F'][i], tab['rho'][i]):
                f.write('\n'.join('%.17g' % x for x in arr) + '\n')
        for arr in tab['rphi']:
            f.write('\n'.join('%.17g' % x for x in arr) + '\n')


def read_funcfl(fn):
    """DYNAMO funcfl layout as parsed by tabulated_eam_init
    (src/potentials/eam/tabulated_eam.f90:172-197): comment; `Z mass a0 lattice`;
    `nF dF nr dr cutoff`; nF F values, nr Z values, nr rho values."""
    with open(fn) as f:
        comment = f.readline().rstrip('\n')
        t = f.readline().split()
        Z, mass, a0, lattice = int(t[0]), float(t[1]), float(t[2]), t[3]
        t = f.readline().replace('D', 'E').split()
        nF, dF, nr, dr, cutoff = int(t[0]), float(t[1]), int(t[2]), float(t[3]), float(t[4])
        rest = np.array(f.read().replace('D', 'E').split(), dtype=np.float64)
    from .elements import chemical_symbols
    return dict(comment=comment, name=chemical_symbols[Z], Znum=Z, mass=mass, a0=a0, lattice=lattice, nF=nF, dF=dF,
                nr=nr, dr=dr, cutoff=cutoff, F=rest[:nF], Z=rest[nF:nF + nr], rho=rest[nF + nr:nF + 2 * nr])
