#!/usr/bin/env python3
"""Golden vectors from the reference's own Fortran kernels, EXECUTED through tests/fortran_subset.py (see
tests/test_func_vs_reference.py): inputs and outputs of the neighbour-list build and of every potential family on
small systems, written to tests/golden/reference_executed.npz.  /root/reference does not travel to the GPU box;
these vectors do, so the GPU kernels are compared with the reference's arithmetic there as well
(tests/test_zz_gpu_reference_vectors.py), and the oracle is compared with them wherever the tests run
(tests/test_reference_vectors.py).

    python tests/golden/make_reference_vectors.py        # needs /root/reference; ~2 min
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import test_func_vs_reference as T            # noqa: E402
from atomistica_b200 import parameters as P   # noqa: E402

OUT = os.path.join(HERE, 'reference_executed.npz')
KEYS = ('epot', 'f', 'wpot', 'epot_per_at', 'wpot_per_at')


def _atoms(store, tag, a, mask=None):
    store[tag + '/symbols'] = np.array(list(a.symbols))
    store[tag + '/positions'] = np.asarray(a.positions, float)
    store[tag + '/cell'] = np.asarray(a.cell, float)
    store[tag + '/pbc'] = np.broadcast_to(np.asarray(a.pbc, bool), (3,)).copy()
    store[tag + '/mask'] = np.zeros(0, np.int32) if mask is None else np.asarray(mask, np.int32)


def _outputs(store, tag, out):
    for k in KEYS:
        store[tag + '/' + k] = np.asarray(out[k], float)
    used = np.flatnonzero(np.abs(out['wpot_per_bond']).reshape(len(out['epot_per_bond']), -1).sum(axis=1)
                          + np.abs(out['epot_per_bond']))
    n = int(used.max()) + 1 if len(used) else 0
    for k in ('epot_per_bond', 'f_per_bond', 'wpot_per_bond'):            # list slots in use only
        store[tag + '/' + k] = np.asarray(out[k], float)[:n]


def _name_of(db):
    return next(n for n in dir(P) if getattr(P, n) is db)


def main():
    store, index = {}, []
    for k, (name, a, cutoff) in enumerate(T._list_cases()):
        seed, last, neighbors, dc, _ = T._reference_neighbor_list(a, cutoff)
        tag = 'nl%d' % k
        _atoms(store, tag, a)
        store[tag + '/cutoff'] = np.float64(cutoff)
        store[tag + '/seed'], store[tag + '/last'] = seed, last
        store[tag + '/neighbors'], store[tag + '/dc'] = neighbors, dc
        index.append((tag, 'nl', name, '', 0, 0))
    from conftest import load_npz
    # the EAM case of test_eam_kernel_executed, through the same code
    setfl = load_npz('cu_mishin1_setfl.npz')
    for k, (out, a, mask) in enumerate(T.run_eam_kernel_cases(setfl)):
        tag = 'eam%d' % k
        _atoms(store, tag, a, mask)
        for key in KEYS:
            store[tag + '/' + key] = np.asarray(out[key], float)
        index.append((tag, 'eam', 'fcc Cu, Cu_mishin1', 'cu_mishin1_setfl.npz', 0, 0))
    rng = np.random.RandomState(77)
    for k, (kind, db, a) in enumerate(T._bop_cases()):
        for m, mask in enumerate((None, (rng.rand(len(a)) > 0.4).astype(np.int32))):
            out, _, _ = T._run_bop_kernel(kind, db, a, mask)
            tag = 'bop%d_%d' % (k, m)
            _atoms(store, tag, a, mask); _outputs(store, tag, out)
            index.append((tag, 'bop', kind, _name_of(db), 0, 0))
    for k, (kind, db, a) in enumerate(T._bop_scr_cases()):
        mask = (rng.rand(len(a)) > 0.4).astype(np.int32) if k == 0 else None
        out, _, _ = T._run_bop_kernel(kind, db, a, mask, screened=True)
        tag = 'bopscr%d' % k
        _atoms(store, tag, a, mask); _outputs(store, tag, out)
        index.append((tag, 'bop', kind, _name_of(db), 1, 0))
    for k, (name, a, dih) in enumerate(T._rebo2_cases()):
        out, _, _ = T._run_rebo2_kernel(a, dih)
        tag = 'rebo2_%d' % k
        _atoms(store, tag, a); _outputs(store, tag, out)
        index.append((tag, 'rebo2', name, '', 0, int(dih)))
    for k, (name, a, dih) in enumerate(T._rebo2_scr_cases()):
        out, _, _ = T._run_rebo2_kernel(a, dih, screened=True)
        tag = 'rebo2scr_%d' % k
        _atoms(store, tag, a); _outputs(store, tag, out)
        index.append((tag, 'rebo2', name, '', 1, int(dih)))
    for k, (name, raw, a, screened) in enumerate(T._juslin_cases()):
        mask = (rng.rand(len(a)) > 0.4).astype(np.int32) if k in (0, 2) else None
        out, _ = T._run_juslin_kernel(raw, a, mask, screened)
        tag = 'juslin%d' % k
        _atoms(store, tag, a, mask); _outputs(store, tag, out)
        index.append((tag, 'juslin', name, _name_of(raw), int(screened), 0))
    store['index'] = np.array(['|'.join(str(x) for x in row) for row in index])
    np.savez_compressed(OUT, **store)
    print('wrote', OUT, len(index), 'cases', os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
