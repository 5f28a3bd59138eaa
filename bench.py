#!/usr/bin/env python3
"""Benchmark of the B200-native Atomistica hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks c2,c1,c3,c4,nl]

Headline workload (BASELINE.json configs[1], "C2"): tabulated alloy EAM (Cu_mishin1.eam.alloy), fcc Cu
40x40x40 = 256,000 atoms per GPU, 300 K Maxwell-Boltzmann velocities (seed 12345), dt = 1 fs, NVE
velocity-Verlet with a Verlet shell, neighbour rebuilds by the reference's rule
2*accum_max_dr >= verlet_shell.  One "step" = one MD step of the whole system.  At N > 1 the global
cell is (40 N) x 40 x 40 fcc cells, slab-decomposed along x (weak scaling).

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; see DESIGN.md section 6.
  value     atom-steps/s with the state resident in HBM (device-resident driver, CUDA events)
  e2e       the same metric through the reference-facing calculator API with HOST buffers
  roofline  dominant kernel (k_eam_force_fast) against the measured HBM copy bandwidth
  parity    N > 1: the decomposed run against the SAME global system on one GPU (rank 0), same number
            of steps: energies and random-weighted force / velocity checksums
  c4        BASELINE configs[3]: Tersoff and Kumagai Si 128^3 cells = 16,777,216 atoms, NVE, strong
            scaling over the N GPUs of the run (its own parity block at N > 1)
  c1, c3, nl_sweep (N = 1)  BASELINE configs[0], [2], [4]
  cpu_baseline  the CPU oracle (restated reference algorithm) on this host's cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
A0, NCELL, TEMP, DT, SKIN = 3.615, 40, 300.0, 1.0, float(os.environ.get('ATX_BENCH_SKIN', '0.5'))
MASS_CU = 63.546
MASS_SI = 28.0855
C4_CELLS = 128
# Verlet shells of C4: the list cutoff r2 + shell stays below the rattled second-neighbour shell (3.84 A), so the
# lists keep ~4 entries per atom while rebuilds are as rare as possible
C4_SKINS = dict(Tersoff=0.6, Kumagai=0.4)
PRIME_MAX = 160          # upper bound of the untimed priming phase (steps)


def load_setfl():
    return dict(np.load(os.path.join(GOLDEN, 'cu_mishin1_setfl.npz'), allow_pickle=False))


def fcc_positions(a0, ncell):
    n = np.broadcast_to(np.asarray(ncell, dtype=int), (3,))
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    ii, jj, kk = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing='ij')
    org = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    return ((org[:, None, :] + basis[None]).reshape(-1, 3) * a0), np.diag(a0 * n.astype(np.float64))


def diamond_positions(a0, ncell):
    n = np.broadcast_to(np.asarray(ncell, dtype=int), (3,))
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    basis = np.concatenate([fcc, fcc + 0.25])
    ii, jj, kk = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing='ij')
    org = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    return ((org[:, None, :] + basis[None]).reshape(-1, 3) * a0), np.diag(a0 * n.astype(np.float64))


def c2_slab(rank):
    """slab `rank` of the weak-scaling C2 system: positions (global frame), velocities, global ids"""
    from atomistica_b200 import md
    pos, cell = fcc_positions(A0, NCELL)
    nat = len(pos)
    pos = pos + rank * cell[0] + 1e-3                      # keep lattice planes off the slab faces
    vel = md.maxwell_boltzmann(np.full(nat, MASS_CU), TEMP, seed=12345 + rank)
    ids = np.arange(nat, dtype=np.int64) + rank * nat
    return pos, vel, ids, cell


def c4_slab(a0, n, rank, world, seed=12345, T=300.0, mass=MASS_SI):
    """diamond Si cells [x0, x1) x n x n of the n^3 supercell, rattled (sigma 0.05 A), global ids and
    Maxwell-Boltzmann velocities.  Random numbers are drawn per x-plane of cells, so the global system is
    the same for every number of ranks."""
    from atomistica_b200.md import ACCEL_CONV, KB
    x0 = (n * rank) // world
    x1 = (n * (rank + 1)) // world
    plane, _ = diamond_positions(a0, (1, n, n))
    kT = KB * T * ACCEL_CONV
    pos = np.empty(((x1 - x0) * len(plane), 3))
    vel = np.empty_like(pos)
    for k, ix in enumerate(range(x0, x1)):
        rng = np.random.RandomState(seed + ix)
        sl = slice(k * len(plane), (k + 1) * len(plane))
        pos[sl] = plane + np.array([ix * a0, 0.0, 0.0]) + 1e-3 + rng.normal(scale=0.05, size=plane.shape)
        vel[sl] = rng.normal(size=plane.shape) * np.sqrt(kT / mass)
    ids = np.arange(len(pos), dtype=np.int64) + 8 * n * n * x0
    return pos, vel, ids


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the GPU work of this process runs (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap,utilization.gpu')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, busy, mx, reasons = [], [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                if float(s[6]) >= 50.0:
                    busy.append(float(s[0]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                      'sw_power_cap'), s[2:6]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        use = busy if busy else sm
        return dict(sm_mhz=float(np.median(use)) if use else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm), samples_under_load=len(busy),
                    window='whole GPU part of this process (all blocks), 50 ms period; median over samples with '
                           'GPU utilisation >= 50 %')


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def flop_counts():
    """instruction-counted FP64 flops per atom of the force kernels (ncu smsp__sass_thread_inst_executed_op_
    d{add,mul,fma}_pred_on of one launch / atoms; written by scripts/summarize_ncu.py)"""
    p = os.path.join(ROOT, 'profiles', 'flops.json')
    return json.load(open(p)) if os.path.exists(p) else {}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle) -- the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------

CPU_FLAGS = ''


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuEamMD:
    """The C2 workload on the host: the oracle's neighbour build and EAM kernel driven by a numpy
    velocity-Verlet with the reference's rebuild rule (standalone/verlet.f90:100-235,
    standalone/neighbors.f90:552-590)."""

    def __init__(self, ncells, threads):
        import oracle
        from atomistica_b200 import md
        global CPU_FLAGS
        self.oracle = oracle
        self.setfl = load_setfl()
        self.eam = oracle.EAM(self.setfl)
        self.r, self.cell = fcc_positions(A0, ncells)
        self.nat = len(self.r)
        self.v = md.maxwell_boltzmann(np.full(self.nat, MASS_CU), TEMP, seed=12345)
        self.eldb = np.full(self.nat, self.eam.eldb(['Cu'])[0], dtype=np.int32)
        CPU_FLAGS = oracle.use_fast(True)
        oracle.set_threads(threads)
        self.acc = md.ACCEL_CONV / MASS_CU
        self.accum = 1e-6
        self.rebuilds = 0
        self.t_build = 0.0
        self.t_force = 0.0
        self.build()
        self.f = self.force()

    def build(self):
        t0 = time.perf_counter()
        self.nl = self.oracle.neighbor_list(self.r, self.cell, [True] * 3, self.eam.cutoff + SKIN, 200)
        self.t_build += time.perf_counter() - t0
        self.rebuilds += 1
        self.accum = 1e-6

    def force(self):
        t0 = time.perf_counter()
        o = self.eam.energy_and_forces(self.r, self.cell, self.nl, self.eldb)
        self.t_force += time.perf_counter() - t0
        self.epot = o['epot']
        return o['f']

    def step(self):
        self.v += 0.5 * self.acc * DT * self.f
        dr = self.v * DT
        self.r += dr
        self.accum += np.sqrt((dr * dr).sum(axis=1).max())
        if 2.0 * self.accum >= SKIN:
            self.build()
        self.f = self.force()
        self.v += 0.5 * self.acc * DT * self.f

    def close(self):
        self.oracle.set_threads(1)
        self.oracle.use_fast(False)


def cpu_best_threads():
    """all host threads unless the serial kernel is faster on this box (cgroup-limited containers)"""
    nt = host_threads()
    if nt == 1:
        return 1
    ts = []
    for th in (1, nt):
        m = CpuEamMD((12, 12, 12), th)
        t0 = time.perf_counter()
        m.force()
        ts.append(time.perf_counter() - t0)
        m.close()
    return nt if ts[1] < ts[0] else 1


def cpu_c2(ncells, steps, warm, threads):
    m = CpuEamMD(ncells, threads)
    try:
        for _ in range(warm):
            m.step()
        r0 = m.rebuilds
        t0 = time.perf_counter()
        for _ in range(steps):
            m.step()
        t = time.perf_counter() - t0
        return dict(nat=m.nat, t=t, rebuilds=m.rebuilds - r0, epot=m.epot)
    finally:
        m.close()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran original cannot be
    built in this image) on the host cores: the SAME workload (256000 atoms per GPU of the run, same
    rebuild rule), real MD steps."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = cpu_best_threads()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # bounded: about 45 ms per step and 256k atoms on 32 threads; cap the whole run at ~3 minutes
    ncells = (NCELL * args.gpus, NCELL, NCELL)
    steps_run = min(steps, max(5, int(200 / (0.06 * args.gpus))))
    warm_run = min(warm, 3)
    res = cpu_c2(ncells, steps_run, warm_run, threads)
    value = res['nat'] * steps_run / res['t']
    sample = ('the full workload: fcc Cu %dx%dx%d cells = %d atoms, %d NVE steps (+%d warm-up) of a numpy '
              'velocity-Verlet around the oracle port (EAM energy/forces + neighbour build, cutoff+%.1f A skin, '
              'rule 2*accum_max_dr >= skin: %d rebuild(s) in the timed steps), %d OpenMP thread(s) of %d host '
              'threads; %s' % (ncells + (res['nat'], steps_run, warm_run, SKIN, res['rebuilds'], threads,
                                         host_threads(), CPU_FLAGS)))
    out = dict(impl='reference', metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=args.gpus,
               steps=steps_run, warmup=warm_run, ms_per_step=res['t'] / steps_run * 1e3, higher_is_better=True,
               scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
               config=config_dict(args.gpus),
               cpu_baseline=dict(value=value, unit='atom-steps/s', cores=threads, kind='port', sample=sample),
               e2e=dict(value=value, unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


def config_dict(ngpu):
    return dict(workload='TabulatedAlloyEAM Cu_mishin1 fcc Cu 40x40x40 (256000 atoms) per GPU, NVE velocity-Verlet, '
                         'dt 1 fs, 300 K, Verlet shell %.2f A' % SKIN,
                atoms_per_gpu=4 * NCELL ** 3, n_gpus=ngpu,
                parallelism='slab domain decomposition x%d along x, peer-to-peer (cudaIpc) halo exchange, NCCL for '
                            'migration' % ngpu if ngpu > 1 else 'single GPU',
                l2_policy='no flush between MD steps (each step consumes the previous one); per-step working set '
                          '(pair list 155 MB + positions/forces) exceeds the 126 MB L2')


# ----------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ----------------------------------------------------------------------------------------------

class Dist:
    """torch.distributed (gloo) for rendezvous, barriers and tiny host reductions only"""

    def __init__(self):
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        self.d = None
        if self.world > 1:
            import torch.distributed as d
            self.d = d
            d.init_process_group(backend='gloo')

    def barrier(self):
        if self.d is not None:
            self.d.barrier()

    def reduce(self, vals, op='sum'):
        vals = np.asarray(vals, dtype=np.float64)
        if self.d is None:
            return vals
        import torch
        t = torch.tensor(vals, dtype=torch.float64)
        self.d.all_reduce(t, op=self.d.ReduceOp.MAX if op == 'max' else self.d.ReduceOp.SUM)
        return t.numpy().copy()

    def close(self):
        if self.d is not None:
            self.d.destroy_process_group()


def prof_read(L, ctx, name):
    import ctypes as C
    tot, cnt = C.c_double(0.0), C.c_longlong(0)
    L.check(L.lib().atx_profile_read(ctx, name.encode(), C.byref(tot), C.byref(cnt)))
    return tot.value, cnt.value


def checksums(ids, v, f):
    """random-weighted sums that any per-atom discrepancy shows up in; weights depend on the global id"""
    w = np.sin(ids.astype(np.float64) * 0.6180339887498949 + 0.3)
    return np.concatenate([(w[:, None] * f).sum(axis=0), [(f * f).sum()], (w[:, None] * v).sum(axis=0),
                           [(v * v).sum()]])


def parity_block(dist, drv, epot, ekin, ref_factory, nsteps_total):
    """decomposed run vs the same global system on rank 0's GPU alone, after the same number of steps"""
    ids, r, v, f = drv.get_state()
    cs = dist.reduce(checksums(ids, v, f))
    out = None
    if dist.rank == 0:
        ref = ref_factory()
        e1, k1 = ref.run(nsteps_total)
        r1, v1, f1 = ref.get_state()
        c1 = checksums(np.arange(len(r1), dtype=np.int64), v1, f1)
        fs, vs = np.sqrt(c1[3]), np.sqrt(c1[7])
        out = dict(reference='the same global system (%d atoms) on one GPU, %d steps from the same initial state'
                             % (len(r1), nsteps_total),
                   epot_per_atom=epot / len(r1), epot_per_atom_single_gpu=e1 / len(r1),
                   epot_rel=abs(epot - e1) / abs(e1), ekin_rel=abs(ekin - k1) / abs(k1),
                   force_checksum_rel=float(np.abs(cs[:3] - c1[:3]).max() / fs),
                   force_norm_rel=float(abs(np.sqrt(cs[3]) - fs) / fs),
                   velocity_checksum_rel=float(np.abs(cs[4:7] - c1[4:7]).max() / vs),
                   rebuilds_single_gpu=ref.stats()['nrebuilds'], tol=1e-9)
        out['ok'] = bool(max(out['epot_rel'], out['ekin_rel'], out['force_checksum_rel'], out['force_norm_rel'],
                             out['velocity_checksum_rel']) < out['tol'])
        del ref
    dist.barrier()
    return out


def prime(drv, nreb=2):
    """untimed: run until `nreb` neighbour rebuilds (with migration under decomposition) have happened, so
    that every buffer has its steady-state size before anything is timed"""
    done = 0
    r0 = drv.stats()['nrebuilds']
    while done < PRIME_MAX and drv.stats()['nrebuilds'] - r0 < nreb:
        drv.run(10)
        done += 10
    return done


def timed_run(dist, L, ctx, drv, steps, profile=False):
    """K steps between barriers, device time of the whole run (CUDA events on the compute stream inside
    the driver, max over ranks).  profile=True additionally brackets every kernel with an event pair
    (atx_profile_*); those runs give the per-kernel times, never the headline numbers -- the event pairs
    cost a few microseconds per launch."""
    dist.barrier()
    L.check(L.lib().atx_ctx_synchronize(ctx))
    L.kernel_launches(reset=True)
    L.check(L.lib().atx_profile_enable(ctx, 1 if profile else 0))
    t0 = time.perf_counter()
    epot, ekin = drv.run(steps)
    wall = time.perf_counter() - t0
    L.check(L.lib().atx_profile_enable(ctx, 0))
    launches = L.kernel_launches()
    ms = float(dist.reduce([drv.stats()['last_run_ms']], 'max')[0])
    dist.barrier()
    return epot, ekin, ms, wall, launches


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------

def block_c2(args, dist, L, ctx):
    from atomistica_b200 import md, native, parallel
    import ctypes as C
    rank, world, local = dist.rank, dist.world, dist.local
    setfl = load_setfl()
    pos, vel, ids, cell = c2_slab(rank)
    nat = len(pos)
    m = np.full(nat, MASS_CU)
    pot = native.TabulatedAlloyEAM(setfl=setfl, device=local)
    if world == 1:
        p = native.from_arrays(np.full(nat, 29, dtype=np.int32), pos, cell, True, device=local)
        nl = native.Neighbors(200, device=local)
        drv = md.VelocityVerlet(pot, p, nl, m, vel, dt=DT, verlet_shell=SKIN)
    else:
        dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=local)
        gcell = cell.copy()
        gcell[0] *= world
        drv = parallel.DDVelocityVerlet(dd, pot, None, [29], gcell, True, ids, np.ones(nat, dtype=np.int32), pos, vel,
                                        m, float(setfl['cutoff']), SKIN, dt=DT)

    steps, warm = args.steps, max(args.warmup, 3)
    primed = prime(drv)
    drv.run(warm)
    st0 = drv.stats()
    epot, ekin, dev_ms, wall, launches = timed_run(dist, L, ctx, drv, steps)
    st = drv.stats()
    rebuilds = st['nrebuilds'] - st0['nrebuilds']
    # the same K steps again with an event pair around every kernel: per-kernel times for the roofline
    _, _, prof_ms, _, _ = timed_run(dist, L, ctx, drv, steps, profile=True)
    prof_rebuilds = drv.stats()['nrebuilds'] - st['nrebuilds']
    st = drv.stats()
    force_ms, force_n = prof_read(L, ctx, 'eam_force')
    dens_ms, dens_n = prof_read(L, ctx, 'eam_density')
    cnt_ms, _ = prof_read(L, ctx, 'nl_pairs_count')
    fill_ms, _ = prof_read(L, ctx, 'nl_pairs_fill')
    dd_prof = None
    if world > 1:
        dd_prof = {k: prof_read(L, ctx, k)[0] for k in ('dd_allreduce', 'dd_halo', 'dd_step')}
        dd_prof['rebuild_host_ms_since_create'] = st['rebuild_host_ms']
        dd_prof['p2p'] = st['p2p']
    z_list = nl.info()['npairs'] / nat if world == 1 else 78.0
    nown, nghost = (nat, 0) if world == 1 else drv.counts()
    value = world * nat * steps / (dev_ms * 1e-3)

    # steady state: a window long enough to hold natural rebuilds at their natural rate
    ss_steps = max(steps, 200)
    st1 = drv.stats()
    e_ss, k_ss, ss_ms, _, _ = timed_run(dist, L, ctx, drv, ss_steps)
    ss_reb = drv.stats()['nrebuilds'] - st1['nrebuilds']
    total_steps = primed + warm + 2 * steps + ss_steps

    parity = None
    if world > 1:
        def ref_factory():
            allp = [c2_slab(k) for k in range(world)]
            gcell = allp[0][3].copy()
            gcell[0] *= world
            gpos = np.concatenate([a[0] for a in allp])
            gvel = np.concatenate([a[1] for a in allp])
            p = native.from_arrays(np.full(len(gpos), 29, dtype=np.int32), gpos, gcell, True, device=local)
            nl1 = native.Neighbors(200, device=local)
            pot1 = native.TabulatedAlloyEAM(setfl=setfl, device=local)
            return md.VelocityVerlet(pot1, p, nl1, np.full(len(gpos), MASS_CU), gvel, dt=DT, verlet_shell=SKIN)
        if not args.no_parity:
            parity = parity_block(dist, drv, e_ss, k_ss, ref_factory, total_steps)
    del drv

    # ---- e2e: reference-facing calculator API, host buffers, copies inside the timed region
    from atomistica_b200 import TabulatedAlloyEAM, structures as S
    calc = TabulatedAlloyEAM(setfl=setfl, device=local, verlet_shell=SKIN)
    pos0, cell0 = fcc_positions(A0, NCELL)
    a2 = S.Atoms(['Cu'] * nat, pos0, cell0, True)
    r = a2.positions
    v = md.maxwell_boltzmann(m, TEMP, seed=12345)
    f = calc.get_forces(a2)          # initialises particles / neighbour list (untimed)
    # priming, as for the device-resident run: untimed calls until the list has been rebuilt once inside the
    # Verlet shell (the first rebuild allocates the fixed-width rows of the single-pass build); the timed
    # calls then see rebuilds at their natural rate
    e2e_primed = 0
    while calc.nl.counters()[0] < 2 and e2e_primed < 120:
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        r += v * DT
        f = calc.get_forces(a2)
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        e2e_primed += 1
    e2e_steps = max(3, min(steps, 20))
    t_api = 0.0
    for k in range(3 + e2e_steps):
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        r += v * DT
        t0 = time.perf_counter()
        f = calc.get_forces(a2)
        dt_call = time.perf_counter() - t0
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        if k >= 3:
            t_api += dt_call
    t_api = float(dist.reduce([t_api], 'max')[0])
    e2e_value = world * nat * e2e_steps / t_api
    del calc
    # the reference's Python host keeps no Verlet shell: the same loop with a list rebuild in every call
    calc = TabulatedAlloyEAM(setfl=setfl, device=local, verlet_shell=0.0)
    f = calc.get_forces(a2)
    n0 = max(3, min(steps, 10))
    t_api0 = 0.0
    for k in range(2 + n0):
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        r += v * DT
        t0 = time.perf_counter()
        f = calc.get_forces(a2)
        dt_call = time.perf_counter() - t0
        v += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        if k >= 2:
            t_api0 += dt_call
    t_api0 = float(dist.reduce([t_api0], 'max')[0])
    e2e_rebuild = dict(value=world * nat * n0 / t_api0, ms_per_call=1e3 * t_api0 / n0, steps=n0,
                       note='verlet_shell = 0: neighbour list rebuilt in every call, the policy of the '
                            "reference's Python host (aseinterface.py:300-333)")
    del calc

    # ---- roofline of the dominant kernel
    peak, peak_src = measured_peaks()
    alg_bytes = nat * (68.0 + 16.0 * z_list)          # B_eam = 68 + 16 z per atom
    # launches that really ran: one force evaluation per step + one per rebuild (the optimistic batches
    # also enqueue steps that exit at once after a rebuild was requested; they carry no time)
    force_n = steps + prof_rebuilds
    force_avg_ms = force_ms / max(force_n, 1)
    achieved = alg_bytes / (force_avg_ms * 1e-3) / 1e9
    fp64 = C.c_double(0.0)
    L.check(L.lib().atx_measure_fp64_peak(ctx, C.byref(fp64)))
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('k_eam_force_dram_bytes_per_launch')
    fl = flop_counts()
    fp64_block = None
    if fl.get('k_eam_force_flops_per_atom'):
        tf = fl['k_eam_force_flops_per_atom'] * nat / (force_avg_ms * 1e-3) / 1e12
        step_flops = (fl['k_eam_force_flops_per_atom'] + fl.get('k_eam_density_flops_per_atom', 0.0)) * nat
        fp64_block = dict(achieved=tf, peak=fp64.value, unit='TFLOP/s', frac=tf / fp64.value,
                          flops_per_atom=fl['k_eam_force_flops_per_atom'],
                          whole_step_frac=step_flops / (dev_ms / steps * 1e-3) / 1e12 / fp64.value,
                          source='ncu instruction counts (profiles/flops.json)')
    out = dict(
        metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=world, steps=steps, warmup=warm,
        ms_per_step=dev_ms / steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
        data='synthetic', config=config_dict(world),
        e2e=dict(value=e2e_value, unit='atom-steps/s', h2d_bytes_per_step=nat * 24, d2h_bytes_per_step=nat * 24 + 80,
                 steps=e2e_steps, ms_per_call=1e3 * t_api / e2e_steps, priming_calls_untimed=e2e_primed,
                 rebuild_every_call=e2e_rebuild,
                 note='calculator API (get_forces returns an array the caller owns, like ASE): host positions in, host '
                      'forces out every call; the positions buffer of the Atoms object is page-locked where it lies '
                      '(same buffer in consecutive calls) and uploaded from directly; neighbour list kept in a %.2f A '
                      'Verlet shell (device-side displacement check every call); at N>1 one calculator instance per '
                      'GPU' % SKIN),
        gpu_launches=launches,
        roofline=dict(bound='hbm', kernel='k_eam_force_fast<4,2,VIRIAL=0,MAP=1>', achieved=achieved, peak=peak,
                      unit='GB/s', frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                      algorithmic_bytes_per_launch=alg_bytes, avg_launch_ms=force_avg_ms, launches=force_n,
                      list_neighbors_per_atom=z_list, share_of_step=force_ms / prof_ms if prof_ms else None,
                      measured_in='an instrumented repeat of the K timed steps (event pair around every kernel; '
                                  '%.4f ms/step against %.4f ms/step un-instrumented)' % (prof_ms / steps, dev_ms / steps),
                      whole_step_frac=alg_bytes / (dev_ms / steps * 1e-3) / 1e9 / peak, fp64=fp64_block),
        kernels_ms=dict(eam_force=force_ms, eam_density=dens_ms, nl_pairs_count=cnt_ms, nl_pairs_fill=fill_ms,
                        total_device=prof_ms, rebuilds=prof_rebuilds, dd=dd_prof,
                        note='instrumented repeat of the K steps'),
        fp64_peak_tflops_measured=fp64.value,
        md=dict(epot=epot, ekin=ekin, rebuilds=rebuilds, wall_s=wall, owned_atoms_rank0=nown,
                ghost_atoms_rank0=nghost, priming_steps_untimed=primed,
                note='priming = untimed steps until two list rebuilds (with migration) have happened, then the W '
                     'warm-up steps, then the K timed steps wherever they fall in the rebuild cycle'),
        steady_state=dict(steps=ss_steps, rebuilds=ss_reb, ms_per_step=ss_ms / ss_steps,
                          value=world * nat * ss_steps / (ss_ms * 1e-3),
                          note='a second timed window long enough to contain rebuilds at their natural rate'),
    )
    if parity is not None:
        out['parity'] = parity
    return out, rebuilds, steps


def block_c4(args, dist, L, ctx):
    """Tersoff / Kumagai Si 128^3 cells (16.8 M atoms), strong scaling over the ranks of this run"""
    from atomistica_b200 import md, native, parallel
    rank, world, local = dist.rank, dist.world, dist.local
    n = args.c4_cells
    peak, _ = measured_peaks()
    fl = flop_counts()
    res = {}
    for kind, a0, rc in (('Tersoff', 5.432, 3.0), ('Kumagai', 5.429, 3.3)):
        if kind not in args.c4_kinds.split(','):
            continue
        C4_SKIN = C4_SKINS[kind]
        pos, v0, ids = c4_slab(a0, n, rank, world)
        cell = np.diag([n * a0] * 3)
        nat, ntot = len(pos), 8 * n ** 3
        m = np.full(nat, MASS_SI)
        pot = getattr(native, kind)(device=local)
        if world == 1:
            p = native.from_arrays(np.full(nat, 14, dtype=np.int32), pos, cell, True, device=local)
            nl = native.Neighbors(20, device=local)
            drv = md.VelocityVerlet(pot, p, nl, m, v0, dt=1.0, verlet_shell=C4_SKIN)
            counts = (nat, 0)
        else:
            dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=local)
            drv = parallel.DDVelocityVerlet(dd, pot, None, [14], cell, True, ids, np.ones(nat, dtype=np.int32), pos, v0,
                                            m, rc, C4_SKIN, dt=1.0, avgn=20)
            counts = drv.counts()
        del pos, v0
        steps = args.c4_steps
        primed = prime(drv)
        drv.run(3)
        st0 = drv.stats()
        epot, ekin, ms, wall, launches = timed_run(dist, L, ctx, drv, steps)
        st = drv.stats()
        epot, ekin, pms, _, _ = timed_run(dist, L, ctx, drv, steps, profile=True)    # per-kernel times
        stp = drv.stats()
        bop_ms, bop_n = prof_read(L, ctx, 'bop_force')
        halo_ms = prof_read(L, ctx, 'dd_halo')[0] if world > 1 else 0.0
        z = nl.info()['npairs'] / nat if world == 1 else None
        blk = dict(workload='%s Si diamond %d^3 cells = %d atoms, rattled 0.05 A, 300 K, NVE dt 1 fs, Verlet shell '
                            '%.1f A' % (kind, n, ntot, C4_SKIN),
                   n_gpus=world, scaling='strong', steps=steps, priming_steps_untimed=primed + 3,
                   ms_per_step=ms / steps, value=ntot * steps / (ms * 1e-3), unit='atom-steps/s',
                   rebuilds=st['nrebuilds'] - st0['nrebuilds'], owned_atoms_rank0=counts[0],
                   ghost_atoms_rank0=counts[1], gpu_launches=launches,
                   bop_force_ms_per_step_rank0=bop_ms / steps, dd_halo_ms_per_step_rank0=halo_ms / steps,
                   instrumented_ms_per_step=pms / steps,
                   epot_per_atom=epot / ntot)
        blk['scopes_ms_per_step_rank0'] = {k: prof_read(L, ctx, k)[0] / steps for k in (
            'bop_force', 'bop_gather', 'dd_drift', 'dd_halo', 'dd_refresh', 'dd_kick', 'nl_update', 'nl_pairs_count',
            'nl_pairs_fill', 'nl_reverse_index')}
        if world > 1:
            blk['p2p'] = st['p2p']
            nreb = max(stp['nrebuilds'] - st0['nrebuilds'], 1)
            blk['rebuild_host_ms_per_rebuild'] = [(a - b) / nreb for a, b in zip(stp['rebuild_host_ms'],
                                                                                 st0['rebuild_host_ms'])]
            blk['rebuild_phases'] = ('migration select + counts, -, migration transfer, ghost select + counts, -, '
                                     'ghost transfer, local list + renumbering, first force evaluation; host wall time, '
                                     'exact per phase only with ATX_DD_PROFILE=1')
        if world == 1:
            alg = nat * (68.0 + 16.0 * z)
            bop_n = steps + (stp['nrebuilds'] - st['nrebuilds'])    # executed evaluations (see block_c2)
            avg = bop_ms / max(bop_n, 1)
            rf = dict(kernel='k_bop_center<%s>' % kind, avg_launch_ms=avg, launches=bop_n,
                      list_neighbors_per_atom=z,
                      hbm=dict(achieved=alg / (avg * 1e-3) / 1e9, peak=peak, unit='GB/s',
                               frac=alg / (avg * 1e-3) / 1e9 / peak, algorithmic_bytes_per_launch=alg))
            fpa = fl.get('k_bop_center_%s_flops_per_atom' % kind)
            if fpa:
                import ctypes as C
                fp64 = C.c_double(0.0)
                L.check(L.lib().atx_measure_fp64_peak(ctx, C.byref(fp64)))
                tf = fpa * nat / (avg * 1e-3) / 1e12
                rf['fp64'] = dict(achieved=tf, peak=fp64.value, unit='TFLOP/s', frac=tf / fp64.value,
                                  flops_per_atom=fpa, source='ncu instruction counts (profiles/flops.json)')
            blk['roofline'] = rf
        else:
            # per-GPU FP64 figure of rank 0: only its OWNED atoms are counted as work (the inner ghost centres it
            # evaluates as well -- +8 % at N=8 -- are not), so the fraction is a lower bound
            fpa = fl.get('k_bop_center_%s_flops_per_atom' % kind)
            if fpa and bop_ms > 0:
                import ctypes as C
                fp64 = C.c_double(0.0)
                L.check(L.lib().atx_measure_fp64_peak(ctx, C.byref(fp64)))
                bop_n = steps + (stp['nrebuilds'] - st['nrebuilds'])
                avg = bop_ms / max(bop_n, 1)
                tf = fpa * counts[0] / (avg * 1e-3) / 1e12
                blk['roofline'] = dict(kernel='k_bop_center<%s>' % kind, avg_launch_ms=avg, launches=bop_n, rank=0,
                                       fp64=dict(achieved=tf, peak=fp64.value, unit='TFLOP/s', frac=tf / fp64.value,
                                                 flops_per_atom=fpa, atoms_counted=int(counts[0]),
                                                 source='ncu instruction counts (profiles/flops.json); owned atoms '
                                                        'of rank 0 only'))

            def ref_factory():
                allp = [c4_slab(a0, n, k, world) for k in range(world)]
                gpos = np.concatenate([a[0] for a in allp])
                gvel = np.concatenate([a[1] for a in allp])
                del allp
                p1 = native.from_arrays(np.full(len(gpos), 14, dtype=np.int32), gpos, cell, True, device=local)
                nl1 = native.Neighbors(20, device=local)
                pot1 = getattr(native, kind)(device=local)
                return md.VelocityVerlet(pot1, p1, nl1, np.full(len(gpos), MASS_SI), gvel, dt=1.0,
                                         verlet_shell=C4_SKIN)
            if not args.no_parity:
                blk['parity'] = parity_block(dist, drv, epot, ekin, ref_factory, primed + 3 + 2 * steps)
        res[kind] = blk
        del drv, pot
        dist.barrier()
    return res


def block_c1(args, L, ctx):
    """BASELINE configs[0]: Tersoff Si 8x8x8 single point through the calculator API"""
    from atomistica_b200 import Tersoff, structures as S
    a = S.diamond('Si', 5.432, (8, 8, 8))
    a.positions += 0.1
    a.rattle(0.05, seed=12345)
    calc = Tersoff()
    rng = np.random.RandomState(1)
    ts = []
    for k in range(25):
        a.positions += rng.normal(scale=1e-6, size=a.positions.shape)   # every call rebuilds, like ASE
        t0 = time.perf_counter()
        calc.calculate(a)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[5:]))
    out = dict(workload='Tersoff Si diamond 8x8x8 (4096 atoms) single-point energy/forces/virial through the '
                        'calculator API, host in / host out, neighbour list rebuilt every call',
               atoms=len(a), ms_per_call=t * 1e3, value=len(a) / t, unit='atom-steps/s',
               energy_per_atom=calc.results['energy'] / len(a))
    if not args.no_cpu:
        import oracle
        from atomistica_b200 import parameters as P
        db = P.complete('Tersoff', None)
        par = oracle.bop_params(oracle.TERSOFF, db)
        el = np.full(len(a), db['el'].index('Si') + 1, dtype=np.int32)
        tc = []
        for _ in range(4):
            t0 = time.perf_counter()
            nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, max(db['r2']), 100)
            o = oracle.bop_energy_and_forces(par, a.positions, a.cell, nl, el)
            tc.append(time.perf_counter() - t0)
        out['cpu_baseline'] = dict(value=len(a) / float(np.median(tc[1:])), unit='atom-steps/s', cores=1, kind='port',
                                   sample='the same 4096-atom configuration, oracle neighbour build + Tersoff kernel, '
                                          '1 thread (the reference wheel is single-threaded), gcc -O2')
        out['parity_epot_rel'] = abs(o['epot'] - calc.results['energy']) / abs(o['epot'])
    return out


def block_c3(args, L, ctx):
    """BASELINE configs[2]: REBO2 on the amorphous-carbon fixture replicated 5x5x5"""
    from atomistica_b200 import native, structures as S
    d = dict(np.load(os.path.join(GOLDEN, 'aC.npz'), allow_pickle=False))
    a = S.Atoms([str(s) for s in d['symbols']], d['positions'], d['cell'], True)
    e1 = None
    big = a.repeat(5)
    p = native.from_atoms(big)
    nl = native.Neighbors(50)
    pot = native.Rebo2()
    pot.bind_to(p, nl)
    for _ in range(3):
        e = pot.energy_and_forces(p, nl)[0]
    L.check(L.lib().atx_profile_enable(ctx, 1))
    rep = 10
    t0 = time.perf_counter()
    for _ in range(rep):
        e = pot.energy_and_forces(p, nl)[0]
    t_call = (time.perf_counter() - t0) / rep
    L.check(L.lib().atx_profile_enable(ctx, 0))
    names = ('rebo2_bonds', 'rebo2_force')
    dev = {k: prof_read(L, ctx, k) for k in names}
    dev_ms = {k: v[0] / max(v[1], 1) for k, v in dev.items()}
    kern = sum(dev_ms.values())
    nat = len(big)
    z = nl.info()['npairs'] / nat
    peak, _ = measured_peaks()
    out = dict(workload='REBO2 amorphous carbon, aC fixture replicated 5x5x5 = %d atoms, energy/forces/virial '
                        '(list reused between calls)' % nat,
               atoms=nat, device_ms=dev_ms, kernels_ms_per_call=kern, value=nat / (kern * 1e-3), unit='atom-steps/s',
               e2e=dict(ms_per_call=t_call * 1e3, value=nat / t_call, note='host forces out every call'),
               energy_per_atom=e / nat, pairs_per_atom=z)
    alg = nat * (68.0 + 16.0 * z + 16.0)
    rf = dict(kernel='k_rebo2_force', avg_launch_ms=dev_ms['rebo2_force'],
              hbm=dict(achieved=alg / (kern * 1e-3) / 1e9, peak=peak, frac=alg / (kern * 1e-3) / 1e9 / peak, unit='GB/s'))
    fpa = flop_counts().get('k_rebo2_force_flops_per_atom')
    if fpa:
        import ctypes as C
        fp64 = C.c_double(0.0)
        L.check(L.lib().atx_measure_fp64_peak(ctx, C.byref(fp64)))
        tf = fpa * nat / (dev_ms['rebo2_force'] * 1e-3) / 1e12
        rf['fp64'] = dict(achieved=tf, peak=fp64.value, unit='TFLOP/s', frac=tf / fp64.value, flops_per_atom=fpa,
                          source='ncu instruction counts (profiles/flops.json)')
    out['roofline'] = rf
    if not args.no_cpu:
        import oracle
        rb = oracle.Rebo2()
        kt = rb.ktyp(a.symbols)
        t0 = time.perf_counter()
        onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, 2.0, 50)
        o = rb.energy_and_forces(a.positions, a.cell, onl, kt)
        tc = time.perf_counter() - t0
        out['cpu_baseline'] = dict(value=len(a) / tc, unit='atom-steps/s', cores=1, kind='port',
                                   sample='the 4001-atom fixture (one replica), oracle neighbour build + REBO2 kernel, 1 '
                                          'thread')
        out['parity_epot_per_atom_rel'] = abs(o['epot'] / len(a) - e / nat) / abs(e / nat)
    return out


def block_nl(args, L, ctx):
    """BASELINE configs[4]: neighbour-list rebuild sweep"""
    from atomistica_b200 import native
    peak, _ = measured_peaks()
    out = []
    sizes = [int(float(x)) for x in args.nl_sizes.split(',')]
    for target in sizes:
        for name, cutoffs in (('Si diamond', (3.0, 3.5)), ('Cu random-density', (5.50679, 6.50679))):
            rng = np.random.RandomState(12345)
            if name == 'Si diamond':
                n = max(2, int(round((target / 8.0) ** (1 / 3))))
                pos, cell = diamond_positions(5.432, n)
                Z, sig, avgn = 14, 0.05, 40
            else:
                # fcc density 0.0847 / A^3; heavily rattled lattice (sigma 0.35 A) stands in for uniform-random
                # positions with a hard core
                n = max(2, int(round((target / 4.0) ** (1 / 3))))
                pos, cell = fcc_positions(3.615, n)
                Z, sig, avgn = 29, 0.35, 200
            for lo in range(0, len(pos), 1 << 22):
                pos[lo:lo + (1 << 22)] += rng.normal(scale=sig, size=pos[lo:lo + (1 << 22)].shape)
            nat = len(pos)
            p = native.from_arrays(np.full(nat, Z, dtype=np.int32), pos, cell, True)
            del pos
            p._sync()
            for rc in cutoffs:
                nl = native.Neighbors(avgn)
                nl.request_interaction_range(rc)
                ts = []
                for k in range(5):
                    L.check(L.lib().atx_ctx_synchronize(ctx))
                    t0 = time.perf_counter()
                    nl.rebuild(p)
                    L.check(L.lib().atx_ctx_synchronize(ctx))
                    ts.append(time.perf_counter() - t0)
                t = float(np.median(ts[2:]))
                info = nl.info()
                z = info['npairs'] / nat
                alg = nat * (40.0 + 16.0 * z)
                out.append(dict(system=name, atoms=nat, cutoff=rc, pairs_per_atom=z, ms=t * 1e3,
                                atoms_per_s=nat / t, pairs_per_s=info['npairs'] / t,
                                hbm_frac=alg / t / 1e9 / peak))
                del nl
            del p
    res = dict(workload='full neighbour-list rebuild (binning, counting sort, pair search, CSR) from positions resident '
                        'in HBM; Si diamond rattled 0.05 A; Cu at fcc density rattled 0.35 A',
               algorithmic_bytes_per_atom='40 + 16 z (SURVEY 8d)', peak_gbs=peak, rows=out)
    if not args.no_cpu:
        import oracle
        rows = []
        for name, rc, (pos, cell) in (('Si diamond', 3.0, diamond_positions(5.432, 23)),
                                      ('Cu random-density', 5.50679, fcc_positions(3.615, 29))):
            rng = np.random.RandomState(12345)
            pos = pos + rng.normal(scale=0.05 if name[0] == 'S' else 0.35, size=pos.shape)
            for th in sorted({1, host_threads()}):
                oracle.use_fast(True)
                oracle.set_threads(th)
                try:
                    t0 = time.perf_counter()
                    oracle.neighbor_list(pos, cell, [True] * 3, rc, 200)
                    tc = time.perf_counter() - t0
                finally:
                    oracle.set_threads(1)
                    oracle.use_fast(False)
                rows.append(dict(system=name, atoms=len(pos), cutoff=rc, cores=th, atoms_per_s=len(pos) / tc))
        res['cpu_baseline'] = dict(kind='port', rows=rows)
    return res


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_ours(args):
    if os.environ.get('ATX_BENCH_WATCHDOG'):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['ATX_BENCH_WATCHDOG']), exit=True)
    from atomistica_b200 import _lib as L
    dist = Dist()
    ctx = L.context(dist.local)
    blocks = set(args.blocks.split(','))
    sampler = ClockSampler(dist.local)
    sampler.start()

    out, rebuilds, steps = block_c2(args, dist, L, ctx)
    errors = {}

    def guarded(name, fn):
        try:
            return fn()
        except Exception as ex:       # a secondary block must not take the headline down
            errors[name] = '%s: %s' % (type(ex).__name__, str(ex)[:300])
            return None

    if 'c4' in blocks:
        out['c4'] = guarded('c4', lambda: block_c4(args, dist, L, ctx))
    if dist.world == 1:
        if 'c1' in blocks:
            out['c1'] = guarded('c1', lambda: block_c1(args, L, ctx))
        if 'c3' in blocks:
            out['c3'] = guarded('c3', lambda: block_c3(args, L, ctx))
        if 'nl' in blocks:
            out['nl_sweep'] = guarded('nl', lambda: block_nl(args, L, ctx))
    out['clocks'] = sampler.finish()
    if errors:
        out['block_errors'] = errors
    if dist.rank == 0:
        if dist.world == 1 and not args.no_cpu:
            threads = cpu_best_threads()
            res = cpu_c2((NCELL, NCELL, NCELL), 12, 2, threads)
            # rebuild cost at the GPU run's natural interval: replace the sample's own rebuild share
            out['cpu_baseline'] = dict(
                value=res['nat'] * 12 / res['t'], unit='atom-steps/s', cores=threads, kind='port',
                sample='the full workload: fcc Cu 40^3 cells = %d atoms, 12 NVE steps (+2 warm-up) of a numpy '
                       'velocity-Verlet around the oracle port with %d OpenMP thread(s) of %d host threads (EAM '
                       'energy/forces + neighbour build, cutoff+%.1f A skin, same rebuild rule: %d rebuild(s) in the '
                       'sample); %s' % (res['nat'], threads, host_threads(), SKIN, res['rebuilds'], CPU_FLAGS))
        emit(out)
    dist.close()


_REAL_STDOUT = None


def capture_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the
    first communicator): from here on file descriptor 1 goes to stderr, and the JSON line is written to the
    original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(out):
    line = json.dumps(out)
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.write(line + '\n')
        _REAL_STDOUT.flush()
    else:
        print(line, flush=True)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs')
    ap.add_argument('--blocks', default='c2,c1,c3,c4,nl', help='secondary blocks to run (c2 always runs)')
    ap.add_argument('--c4-cells', type=int, default=C4_CELLS)
    ap.add_argument('--c4-steps', type=int, default=60)
    ap.add_argument('--nl-sizes', default='1e4,1e5,1e6,4e6,1.6e7,6.4e7')
    ap.add_argument('--c4-kinds', default='Tersoff,Kumagai')
    ap.add_argument('--no-parity', action='store_true', help='skip the single-GPU parity runs at N > 1 (diagnostics)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
