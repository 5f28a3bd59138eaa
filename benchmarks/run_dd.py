#!/usr/bin/env python3
"""C4: Tersoff / Kumagai Si diamond n^3 cells, NVE, slab domain decomposition over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29555 benchmarks/run_dd.py --cells 128 --steps 50 [--kind Tersoff]
    python benchmarks/run_dd.py --cells 128 --steps 50          # N = 1 (single-GPU driver)

Strong scaling: the global system is fixed (cells^3 x 8 atoms); every rank generates only the atoms
of its slab.  Rank 0 prints one JSON line per potential.

    ... benchmarks/run_dd.py --kind Rebo2 --cells 5 --steps 50   # C3: a-C fixture replicated 5^3 (500 k atoms)

REBO2 under decomposition uses a 5 x (rc + skin) halo (every rank evaluates all bonds among its local
atoms and keeps the owned share), dt 0.25 fs."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def slab_atoms(a0, n, rank, world, seed=12345, T=300.0, mass=28.0855):
    """diamond Si cells [x0, x1) x n x n of the n^3 supercell, rattled (sigma 0.05 A), global ids and
    Maxwell-Boltzmann velocities.  Random numbers are drawn per x-plane of cells, so the global system
    is the same for every number of ranks and the energies can be compared across N."""
    from atomistica_b200 import structures as S
    x0 = (n * rank) // world
    x1 = (n * (rank + 1)) // world
    plane = S.diamond('Si', a0, (1, n, n)).positions
    from atomistica_b200.md import ACCEL_CONV
    kT = 8.617333262e-5 * T * ACCEL_CONV      # velocities in Angstrom/fs
    pos, vel = [], []
    for ix in range(x0, x1):
        rng = np.random.RandomState(seed + ix)
        pos.append(plane + np.array([ix * a0, 0.0, 0.0]) + 1e-3 + rng.normal(scale=0.05, size=plane.shape))
        vel.append(rng.normal(size=plane.shape) * np.sqrt(kT / mass))
    pos = np.concatenate(pos)
    vel = np.concatenate(vel)
    ids = np.arange(len(pos), dtype=np.int64) + 8 * n * n * x0
    return pos, vel, ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cells', type=int, default=128)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--kind', default='Tersoff,Kumagai')
    ap.add_argument('--skin', type=float, default=0.4)
    ap.add_argument('--warmup', type=int, default=20)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    from atomistica_b200 import _lib as L, md, native, parallel
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.init_process_group(backend='gloo')
    ctx = L.context(local)
    n = args.cells
    for kind in args.kind.split(','):
        from atomistica_b200 import structures as S
        if kind == 'Rebo2':
            # C3: the amorphous-carbon fixture replicated cells^3 times; every rank builds the whole
            # system (0.5 M atoms) and keeps its slab
            d = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'aC.npz'), allow_pickle=False))
            full = S.Atoms([str(x) for x in d['symbols']], d['positions'], d['cell'], True).repeat(n)
            cell, rc, dt, sym, Z, mass = full.cell, 2.0, 0.25, 'C', 6, 12.011
            vall = md.maxwell_boltzmann(np.full(len(full), mass), 300.0, seed=12345)
            owner = parallel.slab_owner(full.positions, cell, full.pbc, world)
            ids = np.where(owner == rank)[0].astype(np.int64)
            pos, v0 = full.positions[ids], vall[ids]
            ntot, label = len(full), 'C3 Rebo2 a-C %dx%dx%d' % (n, n, n)
            del full, vall
        else:
            a0, rc = (5.432, 3.0) if kind == 'Tersoff' else (5.429, 3.3)
            pos, v0, ids = slab_atoms(a0, n, rank, world)
            cell, dt, sym, Z, mass = np.diag([n * a0] * 3), 1.0, 'Si', 14, 28.0855
            ntot, label = 8 * n ** 3, 'C4 %s Si %d^3 cells' % (kind, n)
        nat = len(pos)
        m = np.full(nat, mass)
        pot = getattr(native, kind)(device=local)
        if world == 1:
            at = S.Atoms([sym] * nat, pos, cell, True)
            p = native.from_atoms(at, device=local)
            nl = native.Neighbors(20, device=local)
            drv = md.VelocityVerlet(pot, p, nl, m, v0, dt=dt, verlet_shell=args.skin)
            counts = (nat, 0)
        else:
            dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=local)
            drv = parallel.DDVelocityVerlet(dd, pot, None, [Z], cell, True, ids, np.ones(nat, dtype=np.int32), pos, v0,
                                            m, rc, args.skin, dt=dt, avgn=20)
            counts = drv.counts()
        drv.run(args.warmup)      # long enough to contain a list rebuild + migration (NCCL connects lazily)
        if dist is not None:
            dist.barrier()
        L.check(L.lib().atx_profile_enable(ctx, 1))
        epot, ekin = drv.run(args.steps)
        L.check(L.lib().atx_profile_enable(ctx, 0))
        st = drv.stats()
        ms = st['last_run_ms']
        import ctypes as C

        def prof(name):
            tot, cnt = C.c_double(0.0), C.c_longlong(0)
            L.check(L.lib().atx_profile_read(ctx, name.encode(), C.byref(tot), C.byref(cnt)))
            return tot.value / args.steps
        prof_ms = {k: prof(k) for k in ('bop_force', 'rebo2_force', 'dd_allreduce', 'dd_halo', 'dd_step', 'nl_pairs_count', 'nl_pairs_fill')}
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        if rank == 0:
            print(json.dumps(dict(config='%s (%d atoms) NVE, skin %.1f' % (label, ntot, args.skin),
                                  n_gpus=world, steps=args.steps, ms_per_step=ms / args.steps,
                                  atom_steps_per_s=ntot * args.steps / (ms * 1e-3), owned_rank0=counts[0],
                                  ghosts_rank0=counts[1], rebuilds=st['nrebuilds'], per_step_ms_rank0=prof_ms,
                                  epot_per_atom=epot / ntot, ekin_per_atom=ekin / ntot)), flush=True)
        del drv, pot
        if dist is not None:
            dist.barrier()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
