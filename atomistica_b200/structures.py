"""Synthetic structure builders for tests and bench.py (ASE is not available in this image).

Conventions follow ase.lattice.cubic (used by the reference's tests): cubic
conventional cells replicated `size` times, cell rows are the cell vectors.
"""
import numpy as np

_FCC = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
_DIA = np.concatenate([_FCC, _FCC + 0.25])


class _Symbols(list):
    """list of chemical symbols that invalidates the owner's cached atomic numbers on assignment"""

    def __init__(self, owner, items):
        super().__init__(items)
        self._owner = owner
        owner._numbers = None

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        self._owner._numbers = None


class Atoms:
    """Minimal stand-in for ase.Atoms: positions, numbers/symbols, cell (rows = vectors), pbc."""

    def __init__(self, symbols, positions, cell, pbc=True):
        self.symbols = _Symbols(self, symbols)
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        cell = np.array(cell, dtype=np.float64)
        self.cell = np.diag(cell) if cell.shape == (3,) else cell
        self.pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).copy()
        self.calc = None

    def __len__(self):
        return len(self.positions)

    def copy(self):
        return Atoms(self.symbols, self.positions.copy(), self.cell.copy(), self.pbc.copy())

    def get_volume(self):
        return abs(np.linalg.det(self.cell))

    def get_atomic_numbers(self):
        """cached like ase.Atoms.numbers; invalidated when a symbol is assigned"""
        if self._numbers is None:
            from .elements import atomic_numbers
            self._numbers = np.array([atomic_numbers[s] for s in self.symbols], dtype=np.int32)
        return self._numbers

    def set_cell(self, cell, scale_atoms=False):
        cell = np.array(cell, dtype=np.float64)
        cell = np.diag(cell) if cell.shape == (3,) else cell
        if scale_atoms:
            s = np.linalg.solve(self.cell.T, self.positions.T).T
            self.positions = s @ cell
        self.cell = cell

    def rattle(self, stdev, seed=42):
        rng = np.random.RandomState(seed)
        self.positions = self.positions + rng.normal(scale=stdev, size=self.positions.shape)

    def repeat(self, rep):
        rep = np.broadcast_to(np.asarray(rep, dtype=int), (3,))
        pos, sym = [], []
        for i in range(rep[0]):
            for j in range(rep[1]):
                for k in range(rep[2]):
                    pos.append(self.positions + np.array([i, j, k]) @ self.cell)
                    sym += self.symbols
        return Atoms(sym, np.concatenate(pos), self.cell * rep[:, None], self.pbc)

    # ASE-style calculator access
    def get_potential_energy(self):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)

    def get_stress(self):
        return self.calc.get_stress(self)


def _cubic(basis, symbols, a, size):
    size = np.broadcast_to(np.asarray(size, dtype=int), (3,))
    ii, jj, kk = np.meshgrid(np.arange(size[0]), np.arange(size[1]), np.arange(size[2]), indexing='ij')
    origins = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    pos = (origins[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    sym = list(symbols) * len(origins)
    return Atoms(sym, pos, a * size.astype(np.float64), True)


def diamond(symbol, a, size=(1, 1, 1)):
    return _cubic(_DIA, [symbol] * 8, a, size)


def fcc(symbol, a, size=(1, 1, 1)):
    return _cubic(_FCC, [symbol] * 4, a, size)


_BCC = np.array([[0, 0, 0], [.5, .5, .5]])
_SC = np.array([[0, 0, 0]])


def bcc(symbol, a, size=(1, 1, 1)):
    return _cubic(_BCC, [symbol] * 2, a, size)


def sc(symbol, a, size=(1, 1, 1)):
    return _cubic(_SC, [symbol], a, size)


def b1(symbols, a, size=(1, 1, 1)):
    """rocksalt: first species on the fcc sites, second on the (1/2,0,0)-shifted fcc sites"""
    return _cubic(np.concatenate([_FCC, (_FCC + np.array([.5, 0, 0])) % 1.0]),
                  [symbols[0]] * 4 + [symbols[1]] * 4, a, size)


def b2(symbols, a, size=(1, 1, 1)):
    """CsCl: simple cubic with a two-atom basis"""
    return _cubic(_BCC, [symbols[0], symbols[1]], a, size)


def b3(symbols, a, size=(1, 1, 1)):
    """zincblende: first species on the fcc sites, second on the (1/4,1/4,1/4)-shifted sites"""
    return _cubic(_DIA, [symbols[0]] * 4 + [symbols[1]] * 4, a, size)


def maxwell_boltzmann(masses_amu, T, seed=12345):
    """velocities in Angstrom/fs-free internal units: returns v in sqrt(eV/amu)"""
    kB = 8.617333262e-5
    rng = np.random.RandomState(seed)
    m = np.asarray(masses_amu, dtype=np.float64)
    v = rng.normal(size=(len(m), 3)) * np.sqrt(kB * T / m)[:, None]
    v -= (v * m[:, None]).sum(axis=0) / m.sum()
    return v


def oriented_diamond(symbols, a, directions, size=(1, 1, 1)):
    """Diamond (one symbol) or zincblende ([A, B]: A on the fcc sites, B on the shifted ones) in the
    smallest orthorhombic cell whose axes run along three mutually orthogonal integer directions, like
    ase.lattice.cubic.Diamond(directions=...).  Along a direction [h, k, l] (coprime) the fcc repeat is
    a/2 [h, k, l] when h + k + l is even and a [h, k, l] otherwise."""
    if isinstance(symbols, str):
        symbols = [symbols, symbols]
    D = np.array(directions, dtype=np.int64)
    assert D.shape == (3, 3) and not (D @ D.T - np.diag(np.diag(D @ D.T))).any(), 'directions must be orthogonal'
    V = np.array([(0.5 if d.sum() % 2 == 0 else 1.0) * a * d for d in D])       # rows: cell vectors
    L = np.sqrt((V * V).sum(axis=1))
    n = int(np.ceil(np.abs(V).sum(axis=0).max() / a)) + 1
    g = np.arange(-n, n + 1)
    cells = np.stack(np.meshgrid(g, g, g, indexing='ij'), axis=-1).reshape(-1, 3)
    pos, sym = [], []
    for b in range(4):
        for shift, s in ((0.0, symbols[0]), (0.25, symbols[1])):
            p = (cells + _FCC[b] + shift) * a
            frac = p @ V.T / (L * L)
            keep = np.all((frac > -1e-9) & (frac < 1 - 1e-9), axis=1)
            pos.append(frac[keep] * L)
            sym += [s] * int(keep.sum())
    pos = np.concatenate(pos)
    order = np.lexsort((pos[:, 0], pos[:, 1], pos[:, 2]))
    unit = Atoms([sym[i] for i in order], pos[order], L, True)
    return unit.repeat(size)
