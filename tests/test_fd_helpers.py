"""atomistica_b200.tests (mirror of the reference's atomistica.tests FD helpers) on a numpy
calculator: the helpers have to find exact forces / stresses consistent and wrong ones not."""
import numpy as np

from atomistica_b200 import structures as S
from atomistica_b200.tests import test_forces as fd_forces, test_virial as fd_virial


class MorseCalc:
    """pair potential under minimum image, ASE calculator protocol"""

    def __init__(self, scale_f=1.0):
        self.scale_f = scale_f

    def _eval(self, atoms):
        r = atoms.positions
        cell = atoms.cell
        d = r[:, None, :] - r[None, :, :]
        s = np.linalg.solve(cell.T, d.reshape(-1, 3).T).T
        d = ((s - np.round(s)) @ cell).reshape(d.shape)
        dist = np.sqrt((d ** 2).sum(-1)) + np.eye(len(r)) * 1e9
        x = np.exp(-1.3 * (dist - 2.6))
        e = 0.5 * np.sum(0.4 * (x * x - 2 * x))
        de = 0.4 * (-2.6 * x * x + 2.6 * x)        # dE_pair/ddist
        f = -np.sum((de / dist)[:, :, None] * d, axis=1)
        w = 0.5 * np.einsum('ij,ija,ijb->ab', de / dist, d, d)      # dE/d(strain)
        return e, f, w

    def get_potential_energy(self, atoms):
        return self._eval(atoms)[0]

    def get_forces(self, atoms):
        return self.scale_f * self._eval(atoms)[1]

    def get_stress(self, atoms):
        w = self._eval(atoms)[2] / atoms.get_volume()
        return np.array([w[0, 0], w[1, 1], w[2, 2], w[1, 2], w[0, 2], w[0, 1]])


def _atoms():
    a = S.fcc('Cu', 3.7, (2, 2, 2))
    a.rattle(0.1, seed=1)
    return a


def test_fd_helpers_accept_consistent_calculator():
    a = _atoms()
    a.calc = MorseCalc()
    p0, c0 = a.positions.copy(), a.cell.copy()
    ffd, f0, maxdf = fd_forces(a, dx=1e-5)
    assert ffd.shape == (len(a), 3) and maxdf < 1e-12
    sfd, s0, maxds = fd_virial(a, de=1e-5)
    assert sfd.shape == (6,) and abs(maxds) < 1e-8
    assert np.abs(sfd - s0).max() < 1e-8
    assert np.allclose(a.positions, p0, atol=1e-12) and np.allclose(a.cell, c0, atol=1e-12)    # state restored


def test_fd_helpers_flag_wrong_forces():
    a = _atoms()
    a.calc = MorseCalc(scale_f=1.01)
    assert fd_forces(a, dx=1e-5)[2] > 1e-8
