"""ASE-style calculator front end, mirroring src/python/atomistica/aseinterface.py.

ASE itself is not importable in this image, so `Atomistica` implements the part of the
ase.calculators.calculator.Calculator protocol the reference's wrapper uses
(get_potential_energy / get_forces / get_stress / get_potential_energies / get_stresses on an
atoms object exposing positions, cell, pbc, symbols) with the same update logic
(aseinterface.py:227-352): exact comparison of cell/pbc/positions, re-initialisation when the
number or type of atoms changes, stress = voigt(wpot)/volume.
"""
import numpy as np

from . import _lib as L
from . import native
from .elements import atomic_numbers


class Atomistica:
    potential_class = None
    avgn = 100

    def __init__(self, potentials=None, avgn=None, device=0, verlet_shell=0.0, zero_copy=False,
                 alias_positions=True, **kwargs):
        """verlet_shell > 0 (Angstrom) keeps the neighbour list between calls until an atom has moved
        verlet_shell/2 from where it was at the last build (checked on the device); 0 rebuilds on every
        change of the positions like the reference's Python host.

        get_forces() / results['forces'] is an array nobody else writes to while the caller holds it
        (like the copy ase.calculators.calculator.Calculator.get_property returns): the library writes
        into a page-locked buffer that is only recycled once no outside reference to it (or to a view
        of it) is left; when more than MAX_FORCE_BUFFERS arrays are held at once the result is copied.
        zero_copy=True skips that copy as well (the caller then must not keep more arrays than that).

        alias_positions: when two consecutive calls present the SAME positions buffer (ase.Atoms updates
        its array in place), that buffer is page-locked where it lies and uploaded from directly instead
        of being copied into a staging mirror first; a caller that switches to another array falls back
        to the mirror.  The contents are read during the call only."""
        self.device = device
        import os
        self.alias_positions = bool(alias_positions) and os.environ.get('ATX_ALIAS_POSITIONS', '1') != '0'
        self._pos_seen = None
        self.zero_copy = bool(zero_copy)
        self.verlet_shell = float(verlet_shell)
        self.pots = potentials if potentials is not None else [self.potential_class(device=device, **kwargs)]
        if avgn is not None:
            self.avgn = avgn
        self.particles = None
        self.nl = None
        self._fbuf = None
        self._fcur = 0
        self._store = False
        self.mask = None
        self.results = {}
        self.kwargs = kwargs
        self.compute_epot_per_bond = False
        self.compute_f_per_bond = False
        self.compute_wpot_per_bond = False
        self.epot_per_bond = self.f_per_bond = self.wpot_per_bond = None

    def todict(self):
        return self.kwargs

    # aseinterface.py:227-289
    def initialize(self, atoms):
        if self.mask is not None and len(self.mask) != len(atoms):
            raise RuntimeError('Length of mask array (= {0}) does not equal number of atoms (= {1}).'
                               .format(len(self.mask), len(atoms)))
        if self.particles is not None:
            self.particles.unalias_coordinates()
        self._pos_seen = None
        self.particles = native.Particles(self.device)
        self.particles.allocate(len(atoms))
        self.particles.set_cell(atoms.cell, atoms.pbc)
        self.particles.Z[:] = atoms.get_atomic_numbers()
        self.particles.coordinates[:, :] = atoms.positions
        self.particles.I_changed_positions()
        self.particles.update_elements()
        self.nl = native.Neighbors(self.avgn, self.device)
        if self.verlet_shell > 0.0:
            self.nl.set(verlet_shell=self.verlet_shell)
        for pot in self.pots:
            pot.bind_to(self.particles, self.nl)

    def set_mask(self, mask):
        self.mask = mask

    def set_per_bond(self, epot=None, f=None, wpot=None):
        if epot is not None:
            self.compute_epot_per_bond = epot
        if f is not None:
            self.compute_f_per_bond = f
        if wpot is not None:
            self.compute_wpot_per_bond = wpot

    # aseinterface.py:300-333
    def update(self, atoms):
        Z = atoms.get_atomic_numbers()
        if self.particles is None or len(self.particles.Z) != len(atoms):
            self.initialize(atoms)
        elif np.any(self.particles.Z != Z):
            self.initialize(atoms)
        if np.any(self.particles.cell != atoms.cell) or np.any(self.particles.pbc != atoms.pbc):
            self.particles.set_cell(atoms.cell, atoms.pbc)
        p = self.particles
        new = atoms.positions
        key = None
        if self.alias_positions and isinstance(new, np.ndarray) and new.dtype == np.float64 and \
                new.flags.c_contiguous and new.flags.aligned and new.shape == (len(p), 3):
            key = (new.ctypes.data, new.shape)
        if p._alias is not None:
            if key is not None and key == self._pos_seen:
                p.I_changed_positions()        # the host's own page-locked array: upload, no mirror copy
                return
            p.unalias_coordinates()            # the host switched to another array: back to the mirror
        elif key is not None and key == self._pos_seen:
            # the same buffer twice in a row: from now on it is r_non_cyc itself
            if p.alias_coordinates(new):
                return
            self.alias_positions = False       # cannot be page-locked here: keep the mirror for good
        self._pos_seen = key
        positions = p.coordinates
        # "did anything move?"  In MD every atom moves, so a strided sample answers it for 1/64 of
        # the cost; only when the sample is unchanged is the full comparison needed.
        if np.any(positions[::64] != new[::64]) or np.any(positions != new):
            positions[:, :] = new
            p.I_changed_positions()

    MAX_FORCE_BUFFERS = 8

    def _force_buffer(self, nat):
        """page-locked force buffer nobody else holds.  With a single potential the library STORES
        the forces straight into it (no zeroing, no host accumulation, no copy).  A buffer is reused
        only when neither it nor a view of it is referenced outside this object (reference counts of
        the array and of the base that owns its views), so an array handed out by get_forces() is
        never overwritten while the caller keeps it -- the guarantee a private copy gives, without
        the copy.  When the caller holds every buffer of the pool the result goes into a fresh
        ordinary array.  Returns (array, private)."""
        import sys
        if self._fbuf is None or self._fbuf[0].array.shape[0] != nat:
            self._fbuf = []
        for pa in self._fbuf:
            arr = pa.array
            # attribute + local + argument = 3; the owner of the views: held by `arr` + argument = 2
            if sys.getrefcount(arr) <= 3 and sys.getrefcount(arr.base) <= 2:
                return arr, True
        if len(self._fbuf) < self.MAX_FORCE_BUFFERS:
            self._fbuf.append(L.PinnedArray((nat, 3)))
            return self._fbuf[-1].array, True
        # every pooled buffer is held by the caller: a fresh pageable array (slower transfer, still private)
        return np.zeros((nat, 3)), True

    # aseinterface.py:355-440
    def calculate(self, atoms, properties=('energy',)):
        self.update(atoms)
        epot = 0.0
        nat = len(self.particles)
        self.results = {}          # drops this object's own reference to the previous force array
        forces, private = self._force_buffer(nat)
        store = len(self.pots) == 1
        if store != self._store:
            for pot in self.pots:
                pot.set_store_outputs(store)
            self._store = store
        if not store:
            forces[...] = 0.0
        wpot = np.zeros((3, 3))
        per_at_e = 'energies' in properties
        per_at_w = 'stresses' in properties
        kwargs = dict(epot_per_at=per_at_e, epot_per_bond=self.compute_epot_per_bond,
                      f_per_bond=self.compute_f_per_bond, wpot_per_at=per_at_w,
                      wpot_per_bond=self.compute_wpot_per_bond)
        if self.mask is not None:
            kwargs['mask'] = self.mask
        epa = wpa = None
        for pot in self.pots:
            _e, _f, _w, _epa, self.epot_per_bond, self.f_per_bond, _wpa, self.wpot_per_bond = \
                pot.energy_and_forces(self.particles, self.nl, forces=forces, **kwargs)
            epot += _e
            wpot += _w
            if _epa is not None:
                epa = _epa if epa is None else epa + _epa
            if _wpa is not None:
                wpa = _wpa if wpa is None else wpa + _wpa
        volume = atoms.get_volume()
        self.results = dict(energy=epot, free_energy=epot, forces=forces if (private or self.zero_copy) else forces.copy(),
                            wpot=wpot)
        self.results['stress'] = np.array([wpot[0, 0], wpot[1, 1], wpot[2, 2], (wpot[1, 2] + wpot[2, 1]) / 2,
                                           (wpot[0, 2] + wpot[2, 0]) / 2, (wpot[0, 1] + wpot[1, 0]) / 2]) / volume
        if per_at_e:
            self.results['energies'] = epa
        if per_at_w:
            self.results['stresses'] = np.transpose([wpa[:, 0, 0], wpa[:, 1, 1], wpa[:, 2, 2],
                                                     (wpa[:, 1, 2] + wpa[:, 2, 1]) / 2,
                                                     (wpa[:, 0, 2] + wpa[:, 2, 0]) / 2,
                                                     (wpa[:, 0, 1] + wpa[:, 1, 0]) / 2])
            # aseinterface.py:438-446: Voigt-ordered wpot_per_at, NOT divided by the volume
        return self.results

    def get_potential_energy(self, atoms):
        return self.calculate(atoms)['energy']

    def get_forces(self, atoms):
        return self.calculate(atoms)['forces']

    def get_stress(self, atoms):
        return self.calculate(atoms)['stress']

    def get_potential_energies(self, atoms):
        return self.calculate(atoms, ('energies',))['energies']

    def get_stresses(self, atoms):
        return self.calculate(atoms, ('stresses',))['stresses']

    # aseinterface.py:459-470
    def get_neighbors(self):
        """(i, j, abs_dr) of the current neighbour list, like the reference's calculator"""
        return self.nl.get_neighbors(self.particles)

    def __str__(self):
        return 'Atomistica([' + ','.join(type(pot).__name__ for pot in self.pots) + '])'


def _calculator(cls):
    # aseinterface.py:491-504: classes ending in 'Scr' get avgn = 1000 (longer lists)
    return type(cls.__name__, (Atomistica,), dict(potential_class=cls, __doc__=cls.__doc__,
                                                  avgn=1000 if cls.__name__.endswith('Scr') else 100))


Tersoff = _calculator(native.Tersoff)
Kumagai = _calculator(native.Kumagai)
Brenner = _calculator(native.Brenner)
TersoffScr = _calculator(native.TersoffScr)
KumagaiScr = _calculator(native.KumagaiScr)
BrennerScr = _calculator(native.BrennerScr)
Juslin = _calculator(native.Juslin)
JuslinScr = _calculator(native.JuslinScr)
LJCut = _calculator(native.LJCut)
Harmonic = _calculator(native.Harmonic)
DoubleHarmonic = _calculator(native.DoubleHarmonic)
BornMayer = _calculator(native.BornMayer)
r6 = _calculator(native.r6)
Rebo2 = _calculator(native.Rebo2)
Rebo2Scr = _calculator(native.Rebo2Scr)
TabulatedAlloyEAM = _calculator(native.TabulatedAlloyEAM)
TabulatedEAM = _calculator(native.TabulatedEAM)
