#!/usr/bin/env python3
"""Design model for the round-2 EAM kernels: how would a cluster-pair (tile) list fill up on the C2
system?  No GPU needed.

The measured limiter of the pair-list kernels is the L1 data pipe: per list
entry ~1 wavefront for the position gather + 0.3 for the entry itself, per in-range pair 2 for the
64-byte table record.  A tile list evaluates MI x MJ atom pairs per list entry: the MJ positions of a
j-cluster are one contiguous 32*MJ-byte load shared by all MI i-atoms, and with Newton's third law
inside the tile each table record is fetched once per UNDIRECTED pair.  What it costs is the
padding: pairs of a tile that are out of range.  This script measures that fill ratio and turns it
into wavefronts per atom with the same model that reproduces the measured kernel times.

    python benchmarks/model_cluster_pairs.py [--cells 16] [--skin 0.5]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomistica_b200 import structures as S   # noqa: E402

RC = 5.50679


def _morton(ix, iy, iz, bits=10):
    k = np.zeros(len(ix), dtype=np.int64)
    for b in range(bits):
        k |= ((ix >> b) & 1).astype(np.int64) << (3 * b + 2)
        k |= ((iy >> b) & 1).astype(np.int64) << (3 * b + 1)
        k |= ((iz >> b) & 1).astype(np.int64) << (3 * b)
    return k


def clusters(pos, L, m, cell):
    """atoms binned into cubic cells of edge `cell`, cell-sorted, chunked into clusters of m.
    cell < 0: atoms sorted along a Morton curve over a grid of edge |cell| instead (spatially compact
    clusters, what a tile list needs)"""
    if cell < 0:
        n = np.maximum(1, np.floor(L / -cell)).astype(int)
        idx = np.floor(pos / (L / n)).astype(int) % n
        order = np.argsort(_morton(idx[:, 0], idx[:, 1], idx[:, 2]), kind='stable')
        ncl = (len(pos) + m - 1) // m
        order = np.concatenate([order, np.full(ncl * m - len(pos), -1)])
        return order.reshape(ncl, m)
    n = np.maximum(1, np.floor(L / cell)).astype(int)
    idx = np.floor(pos / (L / n)).astype(int) % n
    key = (idx[:, 0] * n[1] + idx[:, 1]) * n[2] + idx[:, 2]
    order = np.argsort(key, kind='stable')
    ncl = (len(pos) + m - 1) // m
    pad = ncl * m - len(pos)
    order = np.concatenate([order, np.full(pad, -1)])
    return order.reshape(ncl, m)


def tile_stats(pos, L, mi, mj, rlist, rc, cell):
    ci = clusters(pos, L, mi, cell)
    cj = clusters(pos, L, mj, cell)

    def bbox(c):
        p = np.where(c[..., None] >= 0, pos[np.maximum(c, 0)], np.nan)
        lo, hi = np.nanmin(p, axis=1), np.nanmax(p, axis=1)
        return 0.5 * (lo + hi), 0.5 * (hi - lo)
    ci_c, ci_h = bbox(ci)
    cj_c, cj_h = bbox(cj)
    ntile = 0
    npair_list = 0      # atom pairs within the list cutoff (what the pair list stores)
    npair_in = 0        # atom pairs within rc
    rng = np.random.RandomState(0)
    sample = rng.choice(len(ci), min(len(ci), 400), replace=False)
    for a in sample:
        d = cj_c - ci_c[a]
        d -= np.round(d / L) * L
        gap = np.maximum(np.abs(d) - (cj_h + ci_h[a]), 0.0)
        cand = np.nonzero((gap ** 2).sum(1) < rlist ** 2)[0]
        ia = ci[a][ci[a] >= 0]
        for b in cand:
            jb = cj[b][cj[b] >= 0]
            dr = pos[jb][None, :, :] - pos[ia][:, None, :]
            dr -= np.round(dr / L) * L
            d2 = (dr ** 2).sum(-1)
            if ia[0] == jb[0] or set(ia) & set(jb):
                d2 = np.where(ia[:, None] == jb[None, :], 1e9, d2)
            if (d2 < rlist ** 2).any():          # exact pruning of the tile
                ntile += 1
                npair_list += int((d2 < rlist ** 2).sum())
                npair_in += int((d2 < rc ** 2).sum())
    nat_s = sum(int((ci[a] >= 0).sum()) for a in sample)
    return dict(tiles_per_atom=ntile / nat_s, slots_per_atom=ntile * mi * mj / nat_s,
                list_pairs_per_atom=npair_list / nat_s, in_range_per_atom=npair_in / nat_s,
                fill=npair_in / max(ntile * mi * mj, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cells', type=int, default=16)
    ap.add_argument('--skin', type=float, default=0.5)
    args = ap.parse_args()
    a = S.fcc('Cu', 3.615, (args.cells,) * 3)
    a.rattle(0.08, seed=1)       # ~300 K thermal displacements
    pos = a.positions % np.diag(a.cell)
    L = np.diag(a.cell)
    rlist = RC + args.skin
    print('fcc Cu %d^3 = %d atoms, rc %.3f, list cutoff %.3f' % (args.cells, len(a), RC, rlist))
    print('pair list today: 78 entries/atom, 54 in range -> L1 wavefronts/atom (density + force pass):')
    w_now = 78 * (0.3 + 1.0) * 2 + 54 * 1 + 54 * 2
    print('   list+gather 2x%.0f + density table %d + force table %d = %.0f' % (78 * 1.3, 54, 108, w_now))
    print()
    print('%8s %12s %12s %10s %8s   model wavefronts/atom (both passes, Newton-3 inside tiles)' %
          ('MI x MJ', 'tiles/atom', 'slots/atom', 'in range', 'fill'))
    for mi, mj in ((4, 4), (8, 4), (8, 8), (16, 4)):
        for cell in (3.615, 2 * 3.615, -1.2, -1.8):
            st = tile_stats(pos, L, mi, mj, rlist, RC, cell)
            # per tile and pass: 1 entry + 1 contiguous j-position load (+ MI rows of i in registers);
            # table records: half of the directed in-range pairs per pass (each undirected pair once),
            # density 1 wavefront, force 2; j-side reduction by shuffles: ~ (3 comps * log2(MI)) / MJ per j atom
            w = st['tiles_per_atom'] * mi * (1 + mj * 32 / 128.0) * 2 / mi * 1.0 \
                + 0.5 * st['in_range_per_atom'] * (1 + 2) \
                + st['tiles_per_atom'] * (1 + 3) * np.log2(mi) * 2 / mi
            fp64 = st['slots_per_atom'] / 78.0      # pair slots a warp steps through per listed pair of today
            print('%8s %12.1f %12.0f %10.1f %7.0f%%   %s %.2f A: %.0f  (%.1fx fewer than today); pair slots %.1fx today' %
                  ('%dx%d' % (mi, mj), st['tiles_per_atom'], st['slots_per_atom'], st['in_range_per_atom'],
                   100 * st['fill'], 'cell' if cell > 0 else 'morton grid', abs(cell), w, w_now / w, fp64))


if __name__ == '__main__':
    main()
