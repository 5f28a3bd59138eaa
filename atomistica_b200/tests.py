"""Finite-difference helpers of the reference's `atomistica.tests` (src/python/atomistica/tests.py:
test_forces :44-80, test_virial :83-119), which its own test-suite imports to check every
calculator.  Same return values; written on the position / cell arrays so that they take an
`ase.Atoms` as well as `atomistica_b200.structures.Atoms` (anything with positions, get_cell /
set_cell(scale_atoms=True), get_volume and a calculator behind get_potential_energy / get_forces /
get_stress).
"""
import numpy as np

__test__ = False        # helper library, not a test module


def test_forces(atoms, dx=1e-6):
    """(ffd, f0, max |ffd - f0|^2 per atom): central differences of the energy against the forces"""
    f0 = np.array(atoms.get_forces(), dtype=np.float64).copy()
    ffd = f0.copy()
    for i in range(len(atoms)):
        for c in range(3):
            r0 = atoms.positions[i, c]
            _set_position(atoms, i, c, r0 - dx)
            e1 = atoms.get_potential_energy()
            _set_position(atoms, i, c, r0 + dx)
            e2 = atoms.get_potential_energy()
            _set_position(atoms, i, c, r0)
            ffd[i, c] = -(e2 - e1) / (2 * dx)
    df = ffd - f0
    return ffd, f0, np.max(np.sum(df * df, axis=1))


def test_virial(atoms, de=1e-6):
    """(sfd, s0, max(sfd - s0)): Voigt stress from the strain derivative of the energy against
    get_stress (xx, yy, zz, yz, xz, xy)"""
    s0 = np.array(atoms.get_stress(), dtype=np.float64).copy()
    V0 = atoms.get_volume()
    c0 = np.array(_get_cell(atoms), dtype=np.float64).copy()
    sfd = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            es = []
            for sgn in (-1, 1):
                eps = np.eye(3)
                eps[i, j] += sgn * de
                atoms.set_cell(np.dot(c0, eps), scale_atoms=True)
                es.append(atoms.get_potential_energy())
            sfd[i, j] = (es[1] - es[0]) / (2 * de)
    atoms.set_cell(c0, scale_atoms=True)
    sfd = np.array([sfd[0, 0], sfd[1, 1], sfd[2, 2], (sfd[1, 2] + sfd[2, 1]) / 2, (sfd[0, 2] + sfd[2, 0]) / 2,
                    (sfd[0, 1] + sfd[1, 0]) / 2]) / V0
    return sfd, s0, np.max(sfd - s0)


def _get_cell(atoms):
    return atoms.get_cell() if hasattr(atoms, 'get_cell') else atoms.cell


def _set_position(atoms, i, c, value):
    # assignment through a fresh array so that calculators which compare against a stored copy of
    # the positions (ASE, aseinterface.Atomistica) see the change
    p = np.array(atoms.positions, dtype=np.float64)
    p[i, c] = value
    atoms.positions = p
