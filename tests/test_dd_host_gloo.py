"""CPU test of the N>1 decomposition logic with two gloo ranks: every rank builds its slab + halo
with the host mirror of the device-side ghost construction (atomistica_b200.parallel.local_system)
and evaluates its owned atoms with the oracle; the gathered result must equal the undecomposed
system.  This pins the halo width 2*(rc+skin), the periodic images of the ghosts, the local
non-periodic cell and the per-atom energy partition used by csrc/atx_dd.cu."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from atomistica_b200 import parallel, parameters as P, structures as S
    if case == 'tersoff':
        a = S.diamond('Si', 5.432, (8, 2, 2))
        a.rattle(0.08, seed=1)
        rc, skin = 3.0, 0.3
        db = P.complete('Tersoff', None)
        par = oracle.bop_params(oracle.TERSOFF, db)

        def calc(pos, cell, pbc, symbols):
            nl = oracle.neighbor_list(pos, cell, pbc, rc + skin, 100)
            el = np.array([db['el'].index(s) + 1 for s in symbols], dtype=np.int32)
            return oracle.bop_energy_and_forces(par, pos, cell, nl, el, per_at=True)
    else:
        setfl = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'cu_mishin1_setfl.npz'), allow_pickle=False))
        eam = oracle.EAM(setfl)
        a = S.fcc('Cu', 3.615, (16, 3, 3))
        a.rattle(0.05, seed=2)
        # triclinic variant: shear the cell, atoms follow
        cell = a.cell.copy()
        cell[2, 0] = 0.7
        a.set_cell(cell, scale_atoms=True)
        rc, skin = eam.cutoff, 0.3

        def calc(pos, cell, pbc, symbols):
            nl = oracle.neighbor_list(pos, cell, pbc, rc, 300)
            return eam.energy_and_forces(pos, cell, nl, eam.eldb(symbols), per_at=True)
    own, g, gs, lcell, lpbc, lpos = parallel.local_system(a.positions, a.cell, a.pbc, rank, world, rc, skin)
    sym = [a.symbols[i] for i in np.concatenate([own, g])]
    o = calc(lpos, lcell, lpbc, sym)
    n = len(own)
    res = (own, o['f'][:n], o['epot_per_at'][:n], len(g))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        ref = calc(a.positions, a.cell, a.pbc, a.symbols)
        idx = np.concatenate([x[0] for x in gathered])
        f = np.zeros_like(ref['f']); e = np.zeros(len(a))
        f[idx] = np.concatenate([x[1] for x in gathered])
        e[idx] = np.concatenate([x[2] for x in gathered])
        q.put(dict(nown=[len(x[0]) for x in gathered], nghost=[x[3] for x in gathered], nat=len(a),
                   df=float(np.abs(f - ref['f']).max()), de=float(np.abs(e - ref['epot_per_at']).max()),
                   depot=float(abs(e.sum() - ref['epot']))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('case,port', [('tersoff', 29611), ('eam_triclinic', 29612)])
def test_two_rank_decomposition_matches_global(case, port):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(out['nown']) == out['nat'] and min(out['nghost']) > 0
    assert out['df'] < 1e-10 and out['de'] < 1e-10 and out['depot'] < 1e-9
