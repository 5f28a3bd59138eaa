"""Post-processing helpers on the pair arrays of Neighbors.get_neighbors.

Mirrors `_atomistica.pair_distribution`, `angle_distribution` and `bond_angles`
(src/python/c/analysis.c:29-106, :108-206, :208-318; used by examples/ASE/liquid_tools.py).  The
inputs are the reference's: `i` (and `j`) 0-based, sorted by `i`; the "atoms" the averages run over
are the runs of equal `i`.  Host-side numpy; the per-atom histograms are what h2 needs, so everything
is done with one bincount over (run, bin) pairs.
"""
import numpy as np


def _runs(i):
    """run index of every pair and the number of runs (the reference's `nat`)"""
    i = np.asarray(i)
    if len(i) == 0:
        return np.zeros(0, dtype=np.int64), 0
    start = np.concatenate([[True], i[1:] != i[:-1]])
    run = np.cumsum(start) - 1
    return run, int(run[-1]) + 1


def pair_distribution(i, r, nbins, cutoff):
    """(g(r), variance): per-atom histogram of the pair distances averaged over the atoms and divided
    by the shell volume (analysis.c:29-106)"""
    i = np.asarray(i)
    r = np.asarray(r, dtype=np.float64)
    if i.ndim != 1 or r.ndim != 1:
        raise TypeError('First two arguments need to be one-dimensional arrays.')
    if len(i) != len(r):
        raise RuntimeError('First two arguments need to be arrays of identical length.')
    run, nat = _runs(i)
    nat = max(nat, 1)
    b = (nbins * r / cutoff).astype(np.int64)          # C cast: truncation towards zero
    ok = (b >= 0) & (b < nbins)
    cnt = np.bincount(run[ok] * nbins + b[ok], minlength=nat * nbins).reshape(nat, nbins).astype(np.float64)
    h = cnt.sum(axis=0)
    h2 = (cnt * cnt).sum(axis=0)
    edges = np.arange(nbins + 1) * cutoff / nbins
    binvol = 4 * np.pi / 3 * (edges[1:] ** 3 - edges[:-1] ** 3)
    h = h / (nat * binvol)
    h2 = h2 / (nat * binvol * binvol) - h * h
    return h, h2


def _angles(i, r, cutoff):
    """all ordered pairs (p, p2 != p) of list entries of the same atom with both lengths < cutoff:
    run of p, angle between the two vectors"""
    i = np.asarray(i)
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    run, nat = _runs(i)
    n = (r * r).sum(axis=1)
    keep = np.nonzero(n < cutoff * cutoff)[0]
    runs_k = run[keep]
    # pairs inside each run: block-wise outer products
    first = np.searchsorted(runs_k, np.arange(nat), side='left')
    last = np.searchsorted(runs_k, np.arange(nat), side='right')
    size = last - first
    tot = int((size * size).sum())
    if tot == 0:
        return np.zeros(0, np.int64), np.zeros(0), nat
    owner = np.repeat(np.arange(nat), size * size)
    off = np.arange(tot) - np.repeat(np.cumsum(size * size) - size * size, size * size)
    a = first[owner] + off // size[owner]
    b = first[owner] + off % size[owner]
    m = a != b
    a, b, owner = keep[a[m]], keep[b[m]], owner[m]
    cosang = (r[a] * r[b]).sum(axis=1) / np.sqrt(n[a] * n[b])
    with np.errstate(invalid='ignore'):
        ang = np.arccos(cosang)      # like the C code: no clamping, |cos| > 1 by rounding gives nan
    return owner, ang, nat


def angle_distribution(i, j, r, nbins, cutoff):
    """(distribution of the angles between the bonds of an atom, variance), bonds shorter than cutoff
    (analysis.c:108-206; the normalisation counts one more angle than there are, like the reference)"""
    i = np.asarray(i)
    r = np.asarray(r, dtype=np.float64)
    if r.ndim != 2 or r.shape[1] != 3:
        raise TypeError('Third argument needs to be a two-dimensional double array.')
    if len(i) != len(r) or len(np.asarray(j)) != len(r):
        raise RuntimeError('First three arguments need to be arrays of identical length.')
    owner, ang, nat = _angles(i, r, cutoff)
    nat = max(nat, 1)
    nangle = 1 + len(ang)
    b = (nbins * ang / np.pi).astype(np.int64) % nbins
    cnt = np.bincount(owner * nbins + b, minlength=nat * nbins).reshape(nat, nbins).astype(np.float64)
    binvol = np.pi / nbins
    h = cnt.sum(axis=0) / (nangle * binvol)
    h2 = (cnt * cnt).sum(axis=0) / (nangle * binvol * binvol) - h * h
    return h, h2


def bond_angles(moment, nat, i, j, r, cutoff):
    """per atom: mean of angle**moment over the pairs of its bonds shorter than cutoff, 0 without a
    pair (analysis.c:208-318)"""
    i = np.asarray(i)
    r = np.asarray(r, dtype=np.float64)
    if r.ndim != 2 or r.shape[1] != 3:
        raise TypeError('Fifth argument needs to be a two-dimensional double array.')
    if len(i) != len(r) or len(np.asarray(j)) != len(r):
        raise RuntimeError('First three arguments need to be arrays of identical length.')
    m = np.zeros(nat)
    if len(i) == 0:
        return m
    owner, ang, nruns = _angles(i, r, cutoff)
    run, _ = _runs(i)
    atom_of_run = i[np.concatenate([[0], np.nonzero(i[1:] != i[:-1])[0] + 1])]
    s = np.bincount(owner, weights=ang ** moment, minlength=nruns)
    c = np.bincount(owner, minlength=nruns)
    m[atom_of_run] = np.where(c > 0, s / np.maximum(c, 1), 0.0)
    return m
