"""Spatial domain decomposition over the GPUs of one node (one process per GPU).

Host-side mirror of src/standalone/domain_decomposition.f90 for the part that lives on the host:
the initial assignment of atoms to slabs.  Everything that happens per MD step or per neighbour
rebuild (migration, ghost lists, halo exchange, global rebuild flag) runs on the device inside
libatomistica_b200.so over NCCL (csrc/atx_dd.cu).  torch.distributed, when used, only carries the
128-byte NCCL unique id from rank 0 to the other ranks.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import native
from .md import _KIND


def fractional_x(positions, cell):
    """fractional coordinate along the first cell vector: Bbox(1,:) . r"""
    binv = np.linalg.inv(np.asarray(cell, dtype=np.float64).T)   # Bbox
    return np.asarray(positions) @ binv[0]


def slab_owner(positions, cell, pbc, nranks):
    """rank that owns each atom: s in [k/P, (k+1)/P), periodic wrap or open ends"""
    s = fractional_x(positions, cell)
    if np.broadcast_to(pbc, (3,))[0]:
        s = s - np.floor(s)
    k = np.floor(s * nranks).astype(np.int64)
    return np.clip(k, 0, nranks - 1)


def halo_fraction(cell, rc, skin):
    """halo 2*(rc+skin) as a fraction of the cell thickness along the first cell vector"""
    binv = np.linalg.inv(np.asarray(cell, dtype=np.float64).T)
    return 2.0 * (rc + skin) * np.linalg.norm(binv[0])


def local_system(positions, cell, pbc, rank, nranks, rc, skin):
    """numpy mirror of the device-side ghost construction (used by the CPU tests): returns
    (owned indices, ghost indices, ghost image shift along a1, local cell, local pbc, local positions)"""
    cell = np.asarray(cell, dtype=np.float64)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).copy()
    s = fractional_x(positions, cell)
    if pbc[0]:
        wrap0 = -np.floor(s)
        s = s + wrap0
    else:
        wrap0 = np.zeros_like(s)
    owner = np.clip(np.floor(s * nranks).astype(np.int64), 0, nranks - 1)
    own = np.where(owner == rank)[0]
    if nranks == 1:
        return own, np.zeros(0, int), np.zeros(0), cell, pbc, positions[own] + wrap0[own, None] * cell[0]
    h = halo_fraction(cell, rc, skin)
    slo, shi = rank / nranks, (rank + 1) / nranks
    ghosts, shifts = [], []
    for img in ((-1.0, 0.0, 1.0) if pbc[0] else (0.0,)):
        ss = s + img
        m = (ss >= slo - h) & (ss < shi + h) & ~((owner == rank) & (img == 0.0))
        m &= ~((ss >= slo) & (ss < shi))
        ghosts.append(np.where(m)[0])
        shifts.append(np.full(m.sum(), img))
    g = np.concatenate(ghosts)
    gs = np.concatenate(shifts)
    origin = (slo - h) * cell[0]
    pos_own = positions[own] + wrap0[own, None] * cell[0] - origin
    pos_g = positions[g] + (wrap0[g] + gs)[:, None] * cell[0] - origin
    lcell = cell.copy()
    lcell[0] = cell[0] * (1.0 / nranks + 2 * h)
    lpbc = pbc.copy()
    lpbc[0] = False
    return own, g, gs, lcell, lpbc, np.concatenate([pos_own, pos_g])


class DomainDecomposition:
    """NCCL communicator over the slab neighbours.  `exchange_id(bytes_or_None) -> bytes` must return
    rank 0's unique id on every rank (e.g. a torch.distributed / TCPStore broadcast)."""

    def __init__(self, rank, nranks, exchange_id, device=None):
        self.rank, self.nranks = rank, nranks
        self.device = rank if device is None else device
        self._ctx = L.context(self.device)
        buf = C.create_string_buffer(128)
        if rank == 0:
            L.check(L.lib().atx_dd_get_unique_id(buf))
        uid = exchange_id(bytes(buf.raw) if rank == 0 else None)
        self._h = C.c_void_p()
        L.check(L.lib().atx_dd_create(self._ctx, C.c_int(rank), C.c_int(nranks), C.c_char_p(uid), C.byref(self._h)))

    def __del__(self):
        try:
            L.lib().atx_dd_destroy(self._h)
        except Exception:
            pass


def torch_exchange_id(uid):
    """broadcast helper on top of an initialised torch.distributed process group (any backend)"""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        t = torch.tensor(list(uid), dtype=torch.uint8)
    dist.broadcast(t, src=0)
    return bytes(t.tolist())


class DDVelocityVerlet:
    """Domain-decomposed NVE driver: every rank passes the atoms it owns."""

    def __init__(self, dd, pot, symbols_z, el2Z, cell, pbc, ids, el, positions, velocities, masses, rc, skin,
                 dt=1.0, avgn=200):
        self.dd, self.pot = dd, pot
        abox, bbox = native._abox_bbox(cell)
        ipbc = np.ascontiguousarray(np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).astype(np.int32))
        # bind the potential to the element map only (no particles / neighbours objects on this path)
        a = np.array(el2Z, dtype=np.int32)
        if isinstance(pot, native.TabulatedAlloyEAM):
            from .elements import chemical_symbols
            el2db = np.array([pot.db_elements.index(chemical_symbols[z]) + 1
                              if chemical_symbols[z] in pot.db_elements else -1 for z in el2Z], dtype=np.int32)
            L.check(L.lib().atx_eam_bind_to(pot._h, None, None, C.c_int(len(el2db)), L.iptr(el2db)))
        elif isinstance(pot, native.Rebo2):
            L.check(L.lib().atx_rebo2_bind_to(pot._h, None, None, C.c_int(len(a)), L.iptr(a)))
        else:
            L.check(L.lib().atx_bop_bind_to(pot._h, None, None, C.c_int(len(a)), L.iptr(a)))
        kind = [k for c, k in _KIND.items() if isinstance(pot, c)][0]
        n = len(ids)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        el = np.ascontiguousarray(el, dtype=np.int32)
        r = L.as_f64(positions)
        v = L.as_f64(velocities)
        m = L.as_f64(masses)
        self._h = C.c_void_p()
        L.check(L.lib().atx_dd_md_create(
            dd._h, C.c_int(kind), pot._h, L.dptr(abox), L.dptr(bbox), L.iptr(ipbc), C.c_double(rc), C.c_double(skin),
            C.c_int(avgn), C.c_int(n), ids.ctypes.data_as(C.POINTER(C.c_longlong)), L.iptr(el), L.dptr(r), L.dptr(v),
            L.dptr(m), C.c_double(dt), C.byref(self._h)))

    def __del__(self):
        try:
            L.lib().atx_dd_md_destroy(self._h)
        except Exception:
            pass

    def run(self, nsteps):
        epot, ekin = C.c_double(0.0), C.c_double(0.0)
        L.check(L.lib().atx_dd_md_run(self._h, C.c_int(nsteps), C.byref(epot), C.byref(ekin)))
        return epot.value, ekin.value

    def counts(self):
        a, b = C.c_int(0), C.c_int(0)
        L.check(L.lib().atx_dd_md_get_count(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_state(self):
        n = self.counts()[0]
        ids = np.zeros(n, dtype=np.int64)
        r = np.zeros((n, 3)); v = np.zeros((n, 3)); f = np.zeros((n, 3))
        L.check(L.lib().atx_dd_md_get_state(self._h, ids.ctypes.data_as(C.POINTER(C.c_longlong)), L.dptr(r), L.dptr(v),
                                            L.dptr(f)))
        return ids, r, v, f

    def stats(self):
        n, ms = C.c_longlong(0), C.c_double(0.0)
        L.check(L.lib().atx_dd_md_get_stats(self._h, C.byref(n), C.byref(ms)))
        prof = (C.c_double * 8)()
        p2p = C.c_int(0)
        L.check(L.lib().atx_dd_md_get_profile(self._h, prof, C.byref(p2p)))
        return dict(nrebuilds=n.value, last_run_ms=ms.value, rebuild_host_ms=list(prof), p2p=bool(p2p.value))
