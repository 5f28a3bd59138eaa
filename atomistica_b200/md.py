"""Device-resident NVE driver (velocity-Verlet + Verlet-shell neighbour maintenance).

Mirrors the time loop of the reference's standalone MD code
(src/standalone/main.f90:448-488, verlet.f90:100-235, neighbors.f90:552-590): positions,
velocities and forces stay on the GPU between `run` calls.  Units: eV, Angstrom, amu, fs.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import native

ACCEL_CONV = 9.648533212331e-3   # eV/(A amu) -> A/fs^2
KB = 8.617333262e-5              # eV/K

_KIND = {native.TabulatedAlloyEAM: 1, native._Bop: 2, native.Rebo2: 3}   # ATX_POT_* of the C ABI


class VelocityVerlet:
    def __init__(self, pot, particles, nl, masses, velocities=None, dt=1.0, verlet_shell=None):
        """velocities in Angstrom/fs; verlet_shell (skin) in Angstrom"""
        if verlet_shell is not None:
            nl.set(verlet_shell=verlet_shell)
        pot.bind_to(particles, nl)
        particles._sync()
        self.pot, self.particles, self.nl = pot, particles, nl
        self.nat = len(particles)
        m = L.as_f64(np.broadcast_to(np.asarray(masses, dtype=np.float64), (self.nat,)))
        v = None if velocities is None else L.as_f64(velocities)
        self.masses = m.copy()
        self._h = C.c_void_p()
        kind = [k for c, k in _KIND.items() if isinstance(pot, c)][0]
        L.check(L.lib().atx_md_create(particles._ctx, C.c_int(kind), pot._h, particles._h, nl._h, L.dptr(m),
                                      L.dptr(v), C.c_double(dt), C.byref(self._h)))

    def __del__(self):
        try:
            L.lib().atx_md_destroy(self._h)
        except Exception:
            pass

    def run(self, nsteps):
        epot, ekin = C.c_double(0.0), C.c_double(0.0)
        L.check(L.lib().atx_md_run(self._h, C.c_int(nsteps), C.byref(epot), C.byref(ekin)))
        return epot.value, ekin.value

    def get_state(self):
        r = np.zeros((self.nat, 3)); v = np.zeros((self.nat, 3)); f = np.zeros((self.nat, 3))
        L.check(L.lib().atx_md_get_state(self._h, L.dptr(r), L.dptr(v), L.dptr(f)))
        return r, v, f

    def stats(self):
        n, ms = C.c_longlong(0), C.c_double(0.0)
        L.check(L.lib().atx_md_get_stats(self._h, C.byref(n), C.byref(ms)))
        return dict(nrebuilds=n.value, last_run_ms=ms.value)


def maxwell_boltzmann(masses, T, seed=12345):
    """velocities in Angstrom/fs at temperature T (K), zero total momentum"""
    rng = np.random.RandomState(seed)
    m = np.asarray(masses, dtype=np.float64)
    v = rng.normal(size=(len(m), 3)) * np.sqrt(KB * T / m * ACCEL_CONV)[:, None]
    v -= (v * m[:, None]).sum(axis=0) / m.sum()
    return v
