import json, os, sys
import numpy as np
ROOT = '/root/repo'
sys.path.insert(0, ROOT)
import torch.distributed as dist
from atomistica_b200 import md, native, parallel, structures as S
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group(backend='gloo')
a = S.diamond('Si', 5.432, (16, 4, 4)); a.rattle(0.02, seed=3)
m = np.full(len(a), 28.0855); v0 = np.zeros((len(a), 3))
el = np.ones(len(a), dtype=np.int32)
dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=rank)
owner = parallel.slab_owner(a.positions, a.cell, a.pbc, world)
mine = np.where(owner == rank)[0]
pot = native.Tersoff(device=rank)
drv = parallel.DDVelocityVerlet(dd, pot, None, [14], a.cell, a.pbc, mine, el[mine], a.positions[mine], v0[mine], m[mine], 3.0, 0.3, dt=1e-9)
print(rank, 'counts', drv.counts(), flush=True)
ids, r, v, f0 = drv.get_state()
epot, ekin = drv.run(int(os.environ.get('NSTEP', '1')))
ids1, r, v, f = drv.get_state()
assert np.array_equal(ids, ids1)
g = [None]*world
dist.all_gather_object(g, (ids, r, f, f0))
if rank == 0:
    ids = np.concatenate([x[0] for x in g]); o = np.argsort(ids)
    f = np.concatenate([x[2] for x in g])[o]; r = np.concatenate([x[1] for x in g])[o]
    p = native.from_atoms(a, device=0); nl = native.Neighbors(20, device=0); t = native.Tersoff(device=0); t.bind_to(p, nl)
    e, fr = t.energy_and_forces(p, nl)[:2]
    f0 = np.concatenate([x[3] for x in g])[o]
    print('nbad at create', (np.abs(f0 - fr).max(axis=1) > 1e-8).sum())
    err = np.abs(f - fr).max(axis=1)
    bad = np.where(err > 1e-8)[0]
    sx = a.positions[:, 0] / a.cell[0, 0]
    print('epot', epot, e, 'nbad', len(bad))
    print('bad frac x histogram', np.histogram(sx[bad], bins=12, range=(0, 1))[0])
    print('all  frac x histogram', np.histogram(sx, bins=12, range=(0, 1))[0])
    print('owner of bad', np.bincount(owner[bad], minlength=world))
dist.barrier(); dist.destroy_process_group()
