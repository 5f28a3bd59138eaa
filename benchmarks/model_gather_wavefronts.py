#!/usr/bin/env python3
"""L1 wavefront simulator for the EAM pair kernels (no GPU needed).

The pair kernels are bound by the L1 data pipe (profiles/r01_ncu_eam_final.csv: 83-86 % wavefront
utilisation).  This script replays the LOAD INSTRUCTIONS of k_eam_density_fast / k_eam_force_fast
warp by warp on the C2 geometry (fcc Cu, rattled, list cutoff rc + skin) with the device's atom
order and list order, and counts for every warp-wide load the number of distinct 128-byte lines it
touches (the cost model: one wavefront per distinct line per instruction, >= 1).  It then replays
the same work with other lane -> entry mappings and atom orders:

  current      lane l of a 4-lane group takes entries b + (t*4 + l)*2 + {0,1}   (atx_eam.cu today)
  consecutive  lane l takes entries b + t*8 + u*4 + l: the lanes of one instruction read
               neighbouring entries, which point at neighbouring atoms (runs inside a cell)
  lanes=8/16   more lanes per atom with the consecutive mapping
  morton       atoms inside a binning cell sorted along a Morton curve of 1/4-cell sub-cells instead
               of by original index (list entries stay in cell-stencil order, ascending j per cell)

    python benchmarks/model_gather_wavefronts.py [--cells 12] [--skin 0.5]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomistica_b200 import structures as S   # noqa: E402

RC = 5.50679
NR = 5000            # rows of the Cu_mishin1 r-tables (setfl nr), dr = rc / (nr - 1)


def morton3(ix, iy, iz, bits=4):
    k = np.zeros_like(ix)
    for b in range(bits):
        k |= ((ix >> b) & 1) << (3 * b + 2) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b)
    return k


def device_order(pos, L, rlist, in_cell):
    """sorted atom order of atx_neighbors.cu: by binning cell (edge >= list cutoff); inside a cell by
    original index ('index') or along a Morton curve of sub-cells ('morton')"""
    n = np.maximum(1, np.floor(L / rlist)).astype(int)
    edge = L / n
    c = np.floor(pos / edge).astype(int) % n
    cid = (c[:, 0] * n[1] + c[:, 1]) * n[2] + c[:, 2]
    if in_cell == 'index':
        sub = np.arange(len(pos))
    else:
        q = np.floor((pos - c * edge) / (edge / 4)).astype(int).clip(0, 3)
        sub = morton3(q[:, 0], q[:, 1], q[:, 2], 2)
    order = np.lexsort((np.arange(len(pos)), sub, cid))
    return order, n, cid[order]


def build_lists(pos, L, rlist, in_cell):
    order, n, cid = device_order(pos, L, rlist, in_cell)
    p = pos[order]
    nat = len(p)
    start = np.searchsorted(cid, np.arange(n.prod() + 1))
    lists = []
    for s in range(nat):
        c = cid[s]
        cz = c % n[2]; cy = (c // n[2]) % n[1]; cx = c // (n[2] * n[1])
        ent = []
        for x in (-1, 0, 1):
            for y in (-1, 0, 1):
                for z in (-1, 0, 1):
                    cc = (((cx + x) % n[0]) * n[1] + (cy + y) % n[1]) * n[2] + (cz + z) % n[2]
                    t = np.arange(start[cc], start[cc + 1])
                    d = p[t] - p[s]
                    d -= np.round(d / L) * L
                    d2 = (d ** 2).sum(1)
                    m = (d2 < rlist ** 2) & (t != s)
                    ent.append(np.stack([t[m], d2[m]], axis=1))
        lists.append(np.concatenate(ent))
    return p, lists


def lines(addr, lane, group):
    """distinct 128-byte lines per group of `group` consecutive lanes, summed over the groups of the
    warp (group = 32: the whole warp shares lines; group = 128 B / bytes per lane: a wide load is
    processed in as many passes as it takes to move 128 B per pass, lanes share lines only inside a pass)"""
    addr, lane = np.asarray(addr), np.asarray(lane)
    key = (lane // group) * (1 << 40) + addr // 128
    return len(np.unique(key))


def replay(lists, lanes, unroll, consecutive, split, nwarps=300, seed=0):
    """wavefronts per atom for: list-entry loads (8 B), position gathers (32 B), density table rows
    (32 B), force table records (2 x 32 B)"""
    nat = len(lists)
    gpw = 32 // lanes                     # atoms per warp
    rng = np.random.RandomState(seed)
    warps = rng.choice(nat // gpw, min(nwarps, nat // gpw), replace=False)
    seedo = np.concatenate([[0], np.cumsum([len(x) for x in lists])])
    g8, g32 = (16, 4) if split else (32, 32)
    w_list = w_pos = w_rho = w_rec = 0
    natoms = 0
    for wp in warps:
        atoms = range(wp * gpw, wp * gpw + gpw)
        natoms += gpw
        nmax = max(len(lists[s]) for s in atoms)
        per_iter = lanes * unroll
        for t in range((nmax + per_iter - 1) // per_iter):
            for u in range(unroll):
                a_list, a_pos, a_rho, a_rec, ln, ln_in = [], [], [], [], [], []
                for g, s in enumerate(atoms):
                    ent = lists[s]
                    for l in range(lanes):
                        k = t * per_iter + (u * lanes + l if consecutive else l * unroll + u)
                        if k >= len(ent):
                            continue          # the kernel re-reads its own position: same line as pi
                        j, d2 = int(ent[k, 0]), ent[k, 1]
                        a_list.append((seedo[s] + k) * 8)
                        a_pos.append(j * 32)
                        ln.append(g * lanes + l)
                        if d2 < RC * RC:
                            row = int(np.sqrt(d2) / (RC / (NR - 1)))
                            a_rho.append(row * 32)
                            a_rec.append(row * 64)
                            ln_in.append(g * lanes + l)
                if a_list:
                    w_list += lines(a_list, ln, g8)
                    w_pos += lines(a_pos, ln, g32)
                if a_rho:
                    w_rho += lines(a_rho, ln_in, g32)
                    w_rec += 2 * lines(a_rec, ln_in, g32)   # two 32-byte loads, both in the record's line
    return np.array([w_list, w_pos, w_rho, w_rec]) / natoms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cells', type=int, default=12)
    ap.add_argument('--skin', type=float, default=0.5)
    args = ap.parse_args()
    a = S.fcc('Cu', 3.615, (args.cells,) * 3)
    a.rattle(0.08, seed=1)
    L = np.diag(a.cell)
    pos = a.positions % L
    rlist = RC + args.skin
    print('fcc Cu %d^3 = %d atoms, rc %.3f, list cutoff %.3f' % (args.cells, len(a), RC, rlist))
    print('measured (r01): density pass ~160, force pass ~205 wavefronts/atom at 100 %% pipe utilisation')
    for split in (False, True):
        print()
        print('lines shared by %s' % ('the lanes of one 128-byte pass only (quads for 32-byte loads, half-warps '
                                      'for 8-byte loads)' if split else 'the whole warp'))
        print('%-34s %6s %6s %6s %6s | %8s %8s' % ('variant', 'list', 'pos', 'rho', 'rec', 'density', 'force'))
        for in_cell in ('index', 'morton'):
            p, lists = build_lists(pos, L, rlist, in_cell)
            nl = np.mean([len(x) for x in lists])
            for lanes, unroll, cons, name in ((4, 2, False, 'current'), (4, 2, True, 'consecutive'),
                                              (8, 2, True, 'lanes=8 consecutive'),
                                              (16, 2, True, 'lanes=16 consecutive')):
                w = replay(lists, lanes, unroll, cons, split)
                print('%-34s %6.1f %6.1f %6.1f %6.1f | %8.1f %8.1f' %
                      ('%s order, %s' % (in_cell, name), w[0], w[1], w[2], w[3], w[0] + w[1] + w[2],
                       w[0] + w[1] + w[3]))
    print('(%.1f list entries per atom)' % nl)


if __name__ == '__main__':
    main()
