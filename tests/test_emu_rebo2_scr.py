"""Rebo2Scr device functions run on the CPU against the oracle.

The screened-REBO2 kernels of atomistica_b200/csrc/atx_rebo2.cu are thin wrappers around the
per-atom functions of csrc/atx_rebo2_scr.cuh.  That header is written without CUDA runtime types,
so tests/emu/ compiles THE SAME SOURCE with g++ (device_shim.h stands in for double4, __ldg,
atomicAdd, ...) and runs it one atom after the other on a neighbour list in the device format.
The unscreened kernels are covered the same way (csrc/atx_rebo2_atom.cuh, emu_rebo2).
This test checks that serial run against oracle.Rebo2Scr (which the reference's known answers pin,
tests/test_oracle_kat.py): energy, forces, virial, per-atom and per-bond outputs at 1e-10.
It covers the logic of the kernels, not their launch: the GPU parity test is
tests/test_gpu_rebo2_scr.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from atomistica_b200 import _lib as L
from atomistica_b200 import rebo2_tables as T
from atomistica_b200 import structures as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TOL = 1e-10   # relative to the largest magnitude of the compared array


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('emu') / 'librbs_emu.so')
    cmd = ['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC',
           '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(ROOT, 'atomistica_b200', 'csrc'),
           '-o', out, os.path.join(HERE, 'emu', 'rebo2_scr_emu.cpp')]
    subprocess.run(cmd, check=True)
    return C.CDLL(out)


def device_list(nl, nat):
    """oracle (reference-format) list -> device format: CSR seed without terminator slots, int2 entries
    {0-based neighbour, packed shift}; atoms keep their order (order = identity)"""
    cnt = (nl.last[:nat] - nl.seed[:nat] + 1).astype(np.int64)
    seed = np.zeros(nat + 1, dtype=np.int64)
    seed[1:] = np.cumsum(cnt)
    slots = np.concatenate([np.arange(nl.seed[k] - 1, nl.last[k]) for k in range(nat)]).astype(np.int64) \
        if cnt.sum() else np.zeros(0, np.int64)
    j = nl.neighbors[slots].astype(np.int64) - 1
    dc = nl.dc[slots].astype(np.int64) + 128
    assert dc.min() >= 0 and dc.max() < 256
    ent = np.zeros((len(slots), 2), dtype=np.int32)
    ent[:, 0] = j
    ent[:, 1] = dc[:, 0] | (dc[:, 1] << 8) | (dc[:, 2] << 16)
    return seed, ent, slots


def run_emu(emu, a, nss=32, order=None, screened=True, **kwargs):
    if not screened:
        return run_emu_plain(emu, a, order=order, **kwargs)
    rb = oracle.Rebo2Scr(**kwargs)
    nat = len(a)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, rb.cutoff(a.symbols), 1000)
    ref = rb.energy_and_forces(a.positions, a.cell, nl, rb.ktyp(a.symbols), per_at=True, per_bond=True)
    ref['nl'] = nl

    d, tabs, par, keep, scr, sd = T.build_params_scr(kwargs)
    el2typ = np.zeros(32, dtype=np.int32)
    el2typ[1], el2typ[2] = 1, 3          # particle element ids: 1 = C, 2 = H (3 = anything else)
    pos4 = np.zeros((nat, 4))
    pos4[:, :3] = a.positions
    pos4[:, 3] = [1 if s == 'C' else 2 if s == 'H' else 3 for s in a.symbols]
    seed, ent, slots = device_list(nl, nat)
    order = np.arange(nat, dtype=np.int32) if order is None else np.ascontiguousarray(order, dtype=np.int32)
    abox = oracle.abox_from_cell(a.cell)
    nbs = int(max((seed[1:] - seed[:-1]).max(), 1))
    npairs = len(ent)
    sums = np.zeros(10); f = np.zeros((nat, 3)); epa = np.zeros(nat); wpa = np.zeros((nat, 9))
    stats = np.zeros(4, dtype=np.int32)
    epb = np.zeros(npairs + 1); fpb = np.zeros((npairs + 1, 3)); wpb = np.zeros((npairs + 1, 9))
    flag = emu.emu_rebo2_scr(C.byref(par), C.byref(scr), L.iptr(el2typ), C.c_int(nat), C.c_int(nbs), C.c_int(nss),
                             L.dptr(abox), L.dptr(pos4), seed.ctypes.data_as(C.POINTER(C.c_longlong)),
                             L.iptr(ent), L.iptr(order), L.dptr(sums), L.dptr(f), L.dptr(epa), L.dptr(wpa),
                             L.dptr(epb), L.dptr(fpb), L.dptr(wpb), L.iptr(stats))
    out = dict(flag=flag, stats=stats, epot=sums[0], f=f, wpot=sums[1:].reshape(3, 3).T.copy(), epot_per_at=epa,
               wpot_per_at=wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy())
    # per-bond outputs: device slot n is reference slot slots[n]
    for key, arr in (('epot_per_bond', epb), ('f_per_bond', fpb), ('wpot_per_bond', wpb)):
        full = np.zeros((len(ref[key]),) + arr.shape[1:])
        full[slots] = arr[:npairs]
        out[key] = full.reshape(-1, 3, 3).transpose(0, 2, 1).copy() if key == 'wpot_per_bond' else full
    return out, ref


def run_emu_plain(emu, a, order=None, per_bond=False, **kwargs):
    """unscreened Rebo2: rb_bonds_atom / rb_force_atom (bodies of k_rebo2_bonds / k_rebo2_force)"""
    okw = dict(kwargs)
    if 'dihedral' in okw:
        okw['with_dihedral'] = okw.pop('dihedral')
    rb = oracle.Rebo2(**okw)
    nat = len(a)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, rb.cutoff(a.symbols), 200)
    ref = rb.energy_and_forces(a.positions, a.cell, nl, rb.ktyp(a.symbols), per_at=True, per_bond=True)
    ref['nl'] = nl
    d, tabs, par, keep = T.build_params(kwargs)
    el2typ = np.zeros(32, dtype=np.int32)
    el2typ[1], el2typ[2] = 1, 3
    pos4 = np.zeros((nat, 4))
    pos4[:, :3] = a.positions
    pos4[:, 3] = [1 if s == 'C' else 2 if s == 'H' else 3 for s in a.symbols]
    seed, ent, slots = device_list(nl, nat)
    order = np.arange(nat, dtype=np.int32) if order is None else np.ascontiguousarray(order, dtype=np.int32)
    abox = oracle.abox_from_cell(a.cell)
    nbs = int(max((seed[1:] - seed[:-1]).max(), 1))
    npairs = len(ent)
    sums = np.zeros(10); f = np.zeros((nat, 3)); epa = np.zeros(nat); wpa = np.zeros((nat, 9))
    epb = np.zeros(npairs + 1); fpb = np.zeros((npairs + 1, 3)); wpb = np.zeros((npairs + 1, 9))
    flag = emu.emu_rebo2(C.byref(par), L.iptr(el2typ), C.c_int(nat), C.c_int(nbs), L.dptr(abox), L.dptr(pos4),
                         seed.ctypes.data_as(C.POINTER(C.c_longlong)), L.iptr(ent), L.iptr(order), L.dptr(sums),
                         L.dptr(f), L.dptr(epa), L.dptr(wpa), L.dptr(epb), L.dptr(fpb), L.dptr(wpb), None,
                         C.c_int(1 if per_bond else 0))
    out = dict(flag=flag, epot=sums[0], f=f, wpot=sums[1:].reshape(3, 3).T.copy(), epot_per_at=epa,
               wpot_per_at=wpa.reshape(nat, 3, 3).transpose(0, 2, 1).copy())
    for key, arr in (('epot_per_bond', epb), ('f_per_bond', fpb), ('wpot_per_bond', wpb)):
        full = np.zeros((len(ref[key]),) + arr.shape[1:])
        full[slots] = arr[:npairs]
        out[key] = full.reshape(-1, 3, 3).transpose(0, 2, 1).copy() if key == 'wpot_per_bond' else full
    return out, ref


def close(x, y, what):
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    scale = max(np.abs(y).max(), 1e-3) if y.size else 1.0   # floor: forces of a perfect crystal are rounding noise
    assert np.isfinite(x).all(), what
    err = np.abs(x - y).max() / scale if y.size else 0.0
    assert err < TOL, (what, err)


def check(emu, a, **kw):
    out, ref = run_emu(emu, a, **kw)
    assert out['flag'] == 0
    for key in ('epot', 'f', 'wpot', 'epot_per_at', 'wpot_per_at', 'epot_per_bond', 'f_per_bond', 'wpot_per_bond'):
        close(out[key], ref[key], key)
    return out, ref


def test_amorphous_carbon(emu, aC_small):
    out, ref = check(emu, aC_small)
    assert abs(ref['epot']) > 1.0
    assert out['stats'][1] > 0     # partially screened bonds exist
    print('stats', out['stats'])


def test_rattled_diamond_and_compressed(emu):
    for a0, amp, seed in ((3.566, 0.15, 9), (3.3, 0.2, 3), (3.9, 0.25, 4)):
        a = S.diamond('C', a0, (2, 2, 2))
        a.rattle(amp, seed=seed)
        check(emu, a)


def test_hydrocarbon_solid(emu):
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (2, 2, 2))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    check(emu, a)


def test_foreign_element_is_ignored(emu):
    a = S.diamond('C', 3.6, (2, 2, 2))
    a.symbols[5] = 'Si'
    a.symbols[11] = 'H'
    a.rattle(0.1, seed=2)
    out, ref = check(emu, a)
    assert np.abs(out['f'][5]).max() == 0.0


def test_random_gas_many_screened_bonds(emu):
    # low-density random carbon: many bonds in the screened range 2.25 .. 4 A with partial screening
    rng = np.random.RandomState(7)
    n, box = 60, 9.0
    pos = []
    while len(pos) < n:
        p = rng.uniform(0, box, 3)
        if all(np.linalg.norm((p - q + box / 2) % box - box / 2) > 1.25 for q in pos):
            pos.append(p)
    a = S.Atoms(['C'] * n, np.array(pos), [box, box, box], True)
    out, ref = check(emu, a, nss=64)
    assert out['stats'][1] > 20
    print('stats', out['stats'])


def test_cell_sorted_atom_order(emu, aC_small):
    # the device works on cell-SORTED atoms and decides which of i->j / j->i owns a bond in ORIGINAL
    # numbering (order[]), so that per-bond outputs land in the reference's list slot
    a = aC_small
    nat = len(a)
    perm = np.random.RandomState(3).permutation(nat)       # sorted index s holds original atom perm[s]
    b = S.Atoms([a.symbols[i] for i in perm], a.positions[perm], a.cell, True)
    out, _ = run_emu(emu, b, order=perm)
    _, ref = run_emu(emu, a)
    assert out['flag'] == 0
    close(out['epot'], ref['epot'], 'epot')
    close(out['wpot'], ref['wpot'], 'wpot')
    close(out['f'], ref['f'][perm], 'f')
    close(out['epot_per_at'], ref['epot_per_at'][perm], 'epot_per_at')
    # per-bond: slot of (s -> t, dc) in b's list against slot of (perm[s] -> perm[t], dc) in a's list
    def table(nl, n):
        i, j, dc, slots = oracle.pairs(nl, n)
        return {(int(ii), int(jj), tuple(int(x) for x in d)): int(sl) for ii, jj, d, sl in zip(i, j, dc, slots)}
    rb = oracle.Rebo2Scr()
    nlb = oracle.neighbor_list(b.positions, b.cell, b.pbc, rb.cutoff(b.symbols), 1000)
    ta, tb = table(ref['nl'], nat), table(nlb, nat)
    assert len(ta) == len(tb)
    nz = 0
    for (s_, t_, dc), slot_b in tb.items():
        slot_a = ta[(int(perm[s_]), int(perm[t_]), dc)]
        for key in ('epot_per_bond', 'f_per_bond', 'wpot_per_bond'):
            x, y = out[key][slot_b], ref[key][slot_a]
            assert np.abs(np.asarray(x) - np.asarray(y)).max() < 1e-9, (key, s_, t_)
        nz += ref['epot_per_bond'][slot_a] != 0.0
    assert nz > nat


def test_small_cell_images(emu):
    # one conventional cell: neighbours and screening atoms are periodic images of each other
    a = S.diamond('C', 3.566, (1, 1, 1))
    a.rattle(0.1, seed=5)
    check(emu, a)


def test_other_screening_parameters(emu, aC_small):
    check(emu, aC_small, Cmin=0.8, Cmax=2.6, cc_nc_r2=3.4)


def test_screening_table_overflow_is_flagged(emu, aC_small):
    out, ref = run_emu(emu, aC_small, nss=1)
    assert out['flag'] & 2


def test_alt_dihedral(emu, aC_small):
    """the dihedral term of the screened build (ALT_DIHEDRAL, bop_kernel_rebo2.f90:2089-2371): energy, forces,
    virial, per-atom and per-bond outputs incl. the screening-force factors it feeds into loop 3"""
    out, ref = check(emu, aC_small, with_dihedral=True)
    _, ref0 = run_emu(emu, aC_small)
    assert abs(ref['epot'] - ref0['epot']) > 0.1          # the term is active
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (2, 2, 2))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=2)
    check(emu, a, with_dihedral=True)
    for a0, amp, seed in ((3.566, 0.15, 9), (3.3, 0.2, 3)):
        a = S.diamond('C', a0, (2, 2, 2))
        a.rattle(amp, seed=seed)
        check(emu, a, with_dihedral=True)
    a = S.diamond('C', 3.566, (1, 1, 1))                  # bond partners are periodic images of each other
    a.rattle(0.1, seed=5)
    check(emu, a, with_dihedral=True)
    # cell-sorted (permuted) atom order
    nat = len(aC_small)
    perm = np.random.RandomState(3).permutation(nat)
    b = S.Atoms([aC_small.symbols[i] for i in perm], aC_small.positions[perm], aC_small.cell, True)
    out, _ = run_emu(emu, b, order=perm, with_dihedral=True)
    close(out['epot'], ref['epot'], 'epot')
    close(out['f'], ref['f'][perm], 'f')
    close(out['wpot'], ref['wpot'], 'wpot')


# ---- unscreened Rebo2: the same per-atom source that k_rebo2_bonds / k_rebo2_force wrap ------------

def test_plain_rebo2_amorphous_carbon_and_dihedral(emu, aC_small):
    check(emu, aC_small, screened=False)
    check(emu, aC_small, screened=False, dihedral=True)


def test_plain_rebo2_crystals_and_hydrocarbons(emu):
    a = S.diamond('C', 3.566, (2, 2, 2)); a.rattle(0.1, seed=1)
    check(emu, a, screened=False)
    a = S.diamond('C', 3.566, (1, 1, 1)); a.rattle(0.05, seed=2)
    check(emu, a, screened=False)
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (2, 2, 2))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    check(emu, a, screened=False)
    check(emu, a, screened=False, dihedral=True)


def test_plain_rebo2_caller_supplied_list_with_ghosts(emu):
    """LAMMPS-style operation of Rebo2 (k_rebo2_force_roles): owned atoms plus explicit ghost images,
    full lists without periodic shifts for owned and ghost atoms.  Energy and virial count the owned
    ends of every bond, forces are exact on owned atoms when the ghost shell is 5 bond cutoffs deep."""
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.6, (3, 3, 3))
    for i in rng.choice(len(a), len(a) // 4, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=4)
    rb = oracle.Rebo2()
    cutoff = rb.cutoff(a.symbols)
    nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, cutoff, 200)
    ref = rb.energy_and_forces(a.positions, a.cell, nl, rb.ktyp(a.symbols), per_at=True)

    # unfold: ghosts within 5 cutoffs of the box, brute-force lists (with a skin, as a host code has)
    Lbox = np.diag(a.cell)
    w, skin = 5 * cutoff, 0.3
    pos, sym = [a.positions], [list(a.symbols)]
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                if (sx, sy, sz) == (0, 0, 0):
                    continue
                q = a.positions + np.array([sx, sy, sz]) * Lbox
                m = np.all((q > -w) & (q < Lbox + w), axis=1)
                pos.append(q[m]); sym.append([s for s, t in zip(a.symbols, m) if t])
    pos = np.concatenate(pos); sym = sum(sym, [])
    nall, nloc = len(pos), len(a)
    assert w < Lbox.min()
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, 1e9)
    rows = [np.nonzero(d2[i] < (cutoff + skin) ** 2)[0] for i in range(nall)]
    seed = np.zeros(nall + 1, dtype=np.int64)
    seed[1:] = np.cumsum([len(r) for r in rows])
    ent = np.zeros((seed[-1], 2), dtype=np.int32)
    ent[:, 0] = np.concatenate(rows)
    ent[:, 1] = 128 | (128 << 8) | (128 << 16)
    role = np.ones(nall, dtype=np.uint8); role[:nloc] = 2
    pos4 = np.zeros((nall, 4)); pos4[:, :3] = pos
    pos4[:, 3] = [1 if s == 'C' else 2 for s in sym]
    el2typ = np.zeros(32, dtype=np.int32); el2typ[1], el2typ[2] = 1, 3
    d, tabs, par, keep = T.build_params({})
    abox = np.eye(3).ravel() * 1000.0
    sums = np.zeros(10); f = np.zeros((nall, 3)); epa = np.zeros(nall); wpa = np.zeros((nall, 9))
    order = np.arange(nall, dtype=np.int32)
    flag = emu.emu_rebo2(C.byref(par), L.iptr(el2typ), C.c_int(nall), C.c_int(12), L.dptr(abox), L.dptr(pos4),
                         seed.ctypes.data_as(C.POINTER(C.c_longlong)), L.iptr(ent), L.iptr(order), L.dptr(sums),
                         L.dptr(f), L.dptr(epa), L.dptr(wpa), None, None, None,
                         role.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(0))
    assert flag == 0
    close(sums[0], ref['epot'], 'epot')
    close(sums[1:].reshape(3, 3).T, ref['wpot'], 'wpot')
    close(f[:nloc], ref['f'], 'f')
    close(epa[:nloc], ref['epot_per_at'], 'epot_per_at')
    close(wpa[:nloc].reshape(nloc, 3, 3).transpose(0, 2, 1), ref['wpot_per_at'], 'wpot_per_at')
    assert np.abs(f[nloc:]).max() == 0.0 and np.abs(epa[nloc:]).max() == 0.0


def test_reference_fd_configurations(emu):
    """the structures of the reference's test_forces_and_virial.py rows for Rebo2 / Rebo2Scr (shifted,
    then rattled by 0.5 A: up to 14 bonds and 19 screening neighbours per atom)"""
    import test_gpu_forces_and_virial as G
    for row, screened in ((10, False), (11, True)):
        for name, a in G.table()[row][2]:
            a.positions = a.positions + 0.1
            for state in range(2):
                out, ref = check(emu, a, screened=screened)
                a.rattle(0.5, seed=row + 1)


def test_plain_rebo2_one_thread_per_bond(emu, aC_small):
    """k_rebo2_force_bond (ATX_REBO2_PERBOND=1): compacted list of responsible (atom, slot) pairs, one
    bond per thread, same per-atom source"""
    check(emu, aC_small, screened=False, per_bond=True)
    check(emu, aC_small, screened=False, per_bond=True, dihedral=True)
    a = S.diamond('C', 3.566, (1, 1, 1)); a.rattle(0.05, seed=2)       # self images
    check(emu, a, screened=False, per_bond=True)
    rng = np.random.RandomState(1)
    a = S.diamond('C', 3.7, (2, 2, 2))
    for i in rng.choice(len(a), len(a) // 3, replace=False):
        a.symbols[i] = 'H'
    a.rattle(0.1, seed=6)
    check(emu, a, screened=False, per_bond=True)
    perm = np.random.RandomState(3).permutation(len(aC_small))          # sorted order != original order
    b = S.Atoms([aC_small.symbols[i] for i in perm], aC_small.positions[perm], aC_small.cell, True)
    out, _ = run_emu(emu, b, screened=False, per_bond=True, order=perm)
    _, ref = run_emu(emu, aC_small, screened=False)
    close(out['epot'], ref['epot'], 'epot')
    close(out['f'], ref['f'][perm], 'f')
