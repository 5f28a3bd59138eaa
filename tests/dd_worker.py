"""Worker of tests/test_gpu_dd.py (launched with torch.distributed.run, one rank per GPU):
domain-decomposed MD vs the single-GPU driver from the same initial state."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from atomistica_b200 import md, native, parallel, structures as S
    from atomistica_b200.elements import atomic_numbers
    case = sys.argv[1]
    nsteps = int(sys.argv[2])
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group(backend='gloo')
    if case == 'eam':
        setfl = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'cu_mishin1_setfl.npz'), allow_pickle=False))
        a = S.fcc('Cu', 3.615, (24, 6, 6))
        mass, T, rc, skin, dt = 63.546, 1200.0, float(setfl['cutoff']), 0.4, 2.0
        mk = lambda dev: native.TabulatedAlloyEAM(setfl=setfl, device=dev)
    elif case == 'rebo2':
        # halo 5 x (rc + skin) = 11.5 A: slabs of 21 A at 4 ranks
        a = S.diamond('C', 3.566, (24, 3, 3))
        mass, T, rc, skin, dt = 12.011, 1500.0, 2.0, 0.3, 0.25
        mk = lambda dev: native.Rebo2(device=dev)
    else:
        a = S.diamond('Si', 5.432, (16, 4, 4))
        mass, T, rc, skin, dt = 28.0855, 1500.0, 3.0, 0.3, 1.0
        mk = lambda dev: native.Tersoff(device=dev)
    a.rattle(0.02, seed=3)
    m = np.full(len(a), mass)
    v0 = md.maxwell_boltzmann(m, T, seed=5)
    Z = np.array([atomic_numbers[s] for s in a.symbols])
    el2Z = sorted(set(Z.tolist()))
    el = np.array([el2Z.index(z) + 1 for z in Z], dtype=np.int32)

    dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=rank)
    owner = parallel.slab_owner(a.positions, a.cell, a.pbc, world)
    mine = np.where(owner == rank)[0]
    pot = mk(rank)
    drv = parallel.DDVelocityVerlet(dd, pot, Z, el2Z, a.cell, a.pbc, mine, el[mine], a.positions[mine], v0[mine],
                                    m[mine], rc, skin, dt=dt)
    epot, ekin = drv.run(nsteps)
    ids, r, v, f = drv.get_state()
    st = drv.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ids, r, v, f, st['nrebuilds']))
    out = None
    if rank == 0:
        ids = np.concatenate([g[0] for g in gathered])
        order = np.argsort(ids)
        r = np.concatenate([g[1] for g in gathered])[order]
        v = np.concatenate([g[2] for g in gathered])[order]
        f = np.concatenate([g[3] for g in gathered])[order]
        assert np.array_equal(ids[order], np.arange(len(a)))
        # single-GPU reference run
        p = native.from_atoms(a, device=0)
        nl = native.Neighbors(200, device=0)
        ref = md.VelocityVerlet(mk(0), p, nl, m, v0, dt=dt, verlet_shell=skin)
        e1, k1 = ref.run(nsteps)
        r1, v1, f1 = ref.get_state()
        # positions may differ by whole cell vectors (the DD driver wraps along a1)
        d = r - r1
        s = np.linalg.solve(a.cell.T, d.T).T
        d = (s - np.round(s)) @ a.cell
        out = dict(dr=float(np.abs(d).max()), dv=float(np.abs(v - v1).max()),
                   df=float(np.abs(f - f1).max() / max(1.0, np.abs(f1).max())),
                   depot=abs(epot - e1) / abs(e1), dekin=abs(ekin - k1) / abs(k1),
                   rebuilds=[g[4] for g in gathered], rebuilds_ref=ref.stats()['nrebuilds'], natoms=len(a))
        print('DDRESULT ' + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
