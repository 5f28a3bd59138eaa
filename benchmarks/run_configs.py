#!/usr/bin/env python3
"""Measure the BASELINE.json configurations other than the headline one (bench.py covers C2).

    python benchmarks/run_configs.py [C1] [C3] [C4] [C5] [--cpu] [--out profiles/r01_configs.json]

  C1  Tersoff Si diamond 8x8x8 (4096 atoms), single point through the calculator API
  C3  REBO2 amorphous carbon, aC fixture replicated 5x5x5 (500,125 atoms), E/f/virial
  C4  Tersoff / Kumagai Si diamond up to 128^3 cells (16.8 M atoms): single point + NVE steps (1 GPU;
      multi-GPU numbers come from benchmarks/run_dd.py)
  C5  neighbour-list rebuild sweep 1e4 .. 6.4e7 atoms (Si diamond rc 3.0/3.5; random-density Cu rc
      5.507/6.507)
All timings: CUDA events / perf_counter around synchronous C-ABI calls, 3 warm-ups, median of >= 5.
--cpu adds the single-threaded oracle on a bounded sample (cores = 1).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def med(fn, warm=3, rep=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(rep):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def prof_read(L, ctx, name):
    import ctypes as C
    tot, cnt = C.c_double(0.0), C.c_longlong(0)
    L.check(L.lib().atx_profile_read(ctx, name.encode(), C.byref(tot), C.byref(cnt)))
    return tot.value, cnt.value


def device_times(L, ctx, fn, names, rep=5):
    """average device time (ms) of the named kernels over `rep` calls of fn"""
    for _ in range(3):
        fn()
    L.check(L.lib().atx_profile_enable(ctx, 1))
    for _ in range(rep):
        fn()
    L.check(L.lib().atx_profile_enable(ctx, 0))
    return {n: (lambda t: t[0] / max(t[1], 1))(prof_read(L, ctx, n)) for n in names}


def c1(args):
    from atomistica_b200 import Tersoff, _lib as L, structures as S
    a = S.diamond('Si', 5.432, (8, 8, 8))
    a.positions += 0.1
    a.rattle(0.05, seed=12345)
    calc = Tersoff()
    a.calc = calc
    rng = np.random.RandomState(1)

    def call():
        a.positions += rng.normal(scale=1e-6, size=a.positions.shape)   # force a rebuild like ASE MD
        calc.calculate(a)
    t = med(call, rep=20)
    out = dict(config='C1 Tersoff Si 8x8x8 (4096 atoms) single point via calculator API (host in / host out, '
                      'list rebuilt)', atoms=len(a), ms_per_call=t * 1e3, atom_steps_per_s=len(a) / t,
               energy_per_atom=calc.results['energy'] / len(a))
    ctx = L.context(0)
    out['device_ms'] = device_times(L, ctx, call, ['bop_force', 'nl_pairs_count', 'nl_pairs_fill'], rep=20)
    if args.cpu:
        import oracle
        from atomistica_b200 import parameters as P
        db = P.complete('Tersoff', None)
        par = oracle.bop_params(oracle.TERSOFF, db)
        el = np.full(len(a), 2, dtype=np.int32)

        def cpu():
            nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, 3.0, 100)
            oracle.bop_energy_and_forces(par, a.positions, a.cell, nl, el)
        tc = med(cpu, warm=1, rep=5)
        out['cpu_oracle_1core'] = dict(ms_per_call=tc * 1e3, atom_steps_per_s=len(a) / tc)
    return out


def c3(args):
    from atomistica_b200 import _lib as L, native, structures as S
    d = dict(np.load(os.path.join(GOLDEN, 'aC.npz'), allow_pickle=False))
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    rep = 5 if not args.small else 2
    a = a.repeat(rep)
    p = native.from_atoms(a)
    nl = native.Neighbors(50)
    pot = native.Rebo2()
    pot.bind_to(p, nl)
    res = {}

    def full():
        p.I_changed_positions()
        res['e'] = pot.energy_and_forces(p, nl)[0]

    def list_reused():
        res['e'] = pot.energy_and_forces(p, nl)[0]
    t_full = med(full)
    t_reuse = med(list_reused)
    ctx = L.context(0)
    dev = device_times(L, ctx, list_reused, ['rebo2_force'])
    out = dict(config='C3 REBO2 a-C %dx%dx%d (%d atoms) energy/forces/virial' % (rep, rep, rep, len(a)),
               atoms=len(a), ms_host_in_out_with_rebuild=t_full * 1e3, ms_host_out_list_reused=t_reuse * 1e3,
               device_ms=dev, atom_steps_per_s_kernel=len(a) / (dev['rebo2_force'] * 1e-3),
               atom_steps_per_s_e2e=len(a) / t_full, energy_per_atom=res['e'] / len(a),
               pairs_per_atom=nl.info()['npairs'] / len(a))
    if args.cpu:
        import oracle
        b = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
        rb = oracle.Rebo2()
        kt = rb.ktyp(b.symbols)

        def cpu():
            onl = oracle.neighbor_list(b.positions, b.cell, b.pbc, 2.0, 50)
            rb.energy_and_forces(b.positions, b.cell, onl, kt)
        tc = med(cpu, warm=1, rep=3)
        out['cpu_oracle_1core'] = dict(sample='aC fixture, 4001 atoms', atom_steps_per_s=len(b) / tc)
    return out


def c4(args):
    from atomistica_b200 import _lib as L, md, native, structures as S
    out = []
    n = args.c4_cells
    for kind, cls, a0, rc in (('Tersoff', native.Tersoff, 5.432, 3.0), ('Kumagai', native.Kumagai, 5.429, 3.3)):
        a = S.diamond('Si', a0, (n, n, n))
        a.rattle(0.05, seed=12345)
        nat = len(a)
        p = native.from_atoms(a)
        nl = native.Neighbors(20)
        pot = cls()
        m = np.full(nat, 28.0855)
        v0 = md.maxwell_boltzmann(m, 300.0, seed=12345)
        drv = md.VelocityVerlet(pot, p, nl, m, v0, dt=1.0, verlet_shell=0.4)
        drv.run(5)
        ctx = L.context(0)
        L.check(L.lib().atx_profile_enable(ctx, 1))
        steps = args.c4_steps
        e = drv.run(steps)
        L.check(L.lib().atx_profile_enable(ctx, 0))
        st = drv.stats()
        bop_ms, bop_n = prof_read(L, ctx, 'bop_force')
        z = nl.info()['npairs'] / nat
        alg = nat * (68.0 + 16.0 * z)
        out.append(dict(config='C4 %s Si %d^3 cells (%d atoms) NVE %d steps, skin 0.4, 1 GPU' % (kind, n, nat, steps),
                        atoms=nat, ms_per_step=st['last_run_ms'] / steps,
                        atom_steps_per_s=nat * steps / (st['last_run_ms'] * 1e-3), rebuilds=st['nrebuilds'],
                        bop_center_avg_ms=bop_ms / max(bop_n, 1), list_pairs_per_atom=z,
                        hbm_frac_center_kernel=alg / (bop_ms / max(bop_n, 1) * 1e-3) / 1e9 / 6538.3,
                        epot_per_atom=e[0] / nat))
        del drv, pot, nl, p
    return out


def c5(args):
    from atomistica_b200 import _lib as L, native, structures as S
    ctx = L.context(0)
    out = []
    sizes = [int(x) for x in args.c5_sizes.split(',')]
    for target in sizes:
        for name, cutoffs in (('Si diamond', (3.0, 3.5)), ('random Cu', (5.50679, 6.50679))):
            if name == 'Si diamond':
                n = max(2, int(round((target / 8.0) ** (1 / 3))))
                a = S.diamond('Si', 5.432, (n, n, n))
                a.rattle(0.05, seed=12345)
            else:
                # uniform-random Cu at fcc density with a 1.5 A hard core approximated by a rattled lattice
                n = max(2, int(round((target / 4.0) ** (1 / 3))))
                a = S.fcc('Cu', 3.615, (n, n, n))
                a.rattle(0.35, seed=12345)
            nat = len(a)
            p = native.from_atoms(a)
            for rc in cutoffs:
                nl = native.Neighbors(200 if name != 'Si diamond' else 40)
                nl.request_interaction_range(rc)

                def build():
                    nl.rebuild(p)
                p._sync()
                t = med(build, warm=2, rep=5)
                info = nl.info()
                z = info['npairs'] / nat
                alg = nat * (40.0 + 16.0 * z)
                out.append(dict(system=name, atoms=nat, cutoff=rc, pairs_per_atom=z, ms=t * 1e3,
                                atoms_per_s=nat / t, pairs_per_s=info['npairs'] / t,
                                hbm_frac=alg / t / 1e9 / 6538.3))
                del nl
            del p
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('which', nargs='*', default=['C1', 'C3', 'C4', 'C5'])
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--small', action='store_true')
    ap.add_argument('--c4-cells', type=int, default=128)
    ap.add_argument('--c4-steps', type=int, default=50)
    ap.add_argument('--c5-sizes', default='10000,100000,1000000,4000000,16000000')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    res = {}
    for w in args.which:
        res[w] = dict(C1=c1, C3=c3, C4=c4, C5=c5)[w](args)
        print(w, json.dumps(res[w], indent=1), flush=True)
    if args.out:
        json.dump(res, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
