#!/usr/bin/env python3
"""Headline benchmark of the B200-native Atomistica hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): tabulated alloy EAM (Cu_mishin1.eam.alloy), fcc Cu
40x40x40 = 256,000 atoms per GPU, 300 K Maxwell-Boltzmann velocities (seed 12345), dt = 1 fs,
NVE velocity-Verlet with a Verlet shell, neighbour rebuilds by the reference's rule
2*accum_max_dr >= verlet_shell.  One "step" = one MD step of the whole system.

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; see DESIGN.md section 6.
  value    atom-steps/s with the state resident in HBM (device-resident driver, CUDA events)
  e2e      the same metric through the reference-facing calculator API with HOST buffers:
           every step copies the positions host->device and the forces device->host and, like the
           reference's Python host (no skin), rebuilds the neighbour list
  roofline dominant kernel (k_eam_force) against the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (restated reference algorithm) on this host, bounded sample
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
A0, NCELL, TEMP, DT, SKIN = 3.615, 40, 300.0, 1.0, 0.5
MASS_CU = 63.546


def load_setfl():
    return dict(np.load(os.path.join(GOLDEN, 'cu_mishin1_setfl.npz'), allow_pickle=False))


def build_system(ncell=(NCELL, NCELL, NCELL)):
    from atomistica_b200 import md, structures as S
    a = S.fcc('Cu', A0, ncell)
    m = np.full(len(a), MASS_CU)
    v = md.maxwell_boltzmann(m, TEMP, seed=12345)
    return a, m, v


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
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
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                      'sw_power_cap'), s[2:6]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle) -- the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------

CPU_FLAGS = ''


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(ncell=12, nforce=3, threads=1):
    """Bounded sample of the same workload on the host: fcc Cu ncell^3 cells, one neighbour build
    (cutoff + skin) and `nforce` EAM force evaluations of the oracle, both with `threads` OpenMP
    threads (the reference runs this kernel under "!$omp parallel" with thread-local force arrays,
    tabulated_alloy_eam.f90:473-486)."""
    import oracle
    from atomistica_b200 import structures as S
    setfl = load_setfl()
    eam = oracle.EAM(setfl)
    a = S.fcc('Cu', A0, (ncell, ncell, ncell))
    a.rattle(0.05, seed=12345)
    eldb = eam.eldb(a.symbols)
    global CPU_FLAGS
    CPU_FLAGS = oracle.use_fast(True)
    oracle.set_threads(threads)
    try:
        t0 = time.perf_counter()
        nl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff + SKIN, 200)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(nforce):
            eam.energy_and_forces(a.positions, a.cell, nl, eldb)
        t_force = (time.perf_counter() - t0) / nforce
    finally:
        oracle.set_threads(1)
        oracle.use_fast(False)
    return len(a), t_build, t_force


def cpu_best_threads(ncell=12):
    """all host threads unless the serial kernel is faster on this box (cgroup-limited containers)"""
    nt = host_threads()
    cpu_sample(ncell=8, nforce=1)          # builds / loads the oracle library
    if nt == 1:
        return 1
    t1 = cpu_sample(ncell=ncell, nforce=1, threads=1)[2]
    tn = cpu_sample(ncell=ncell, nforce=1, threads=nt)[2]
    return nt if tn < t1 else 1


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran original cannot be
    built in this image) on the host cores, same metric/config, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = cpu_best_threads()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    interval = 33
    ncell = 24
    for _ in range(min(warm, 2)):
        cpu_sample(ncell=ncell, nforce=1, threads=threads)
    t_steps = []
    for _ in range(min(steps, 5)):
        nat, t_build, t_force = cpu_sample(ncell=ncell, nforce=2, threads=threads)
        t_steps.append(t_force + t_build / interval)
    t = float(np.mean(t_steps))
    value = nat / t
    sample = ('fcc Cu %d^3 cells = %d atoms per step (bounded sample of the 256000-atom workload), oracle port, '
              '%d OpenMP thread(s) of %d host threads for the EAM energy/forces and the neighbour build '
              '(cutoff+%.1f A skin, 1 build amortised over %d steps); %s' % (ncell, nat, threads, host_threads(), SKIN, interval, CPU_FLAGS))
    out = dict(impl='reference', metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=args.gpus,
               steps=min(steps, 5), warmup=min(warm, 2), ms_per_step=t * 1e3, higher_is_better=True,
               scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
               config=config_dict(args.gpus),
               cpu_baseline=dict(value=value, unit='atom-steps/s', cores=threads, kind='port', sample=sample),
               e2e=dict(value=value, unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    print(json.dumps(out))


def config_dict(ngpu):
    return dict(workload='TabulatedAlloyEAM Cu_mishin1 fcc Cu 40x40x40 (256000 atoms) per GPU, NVE velocity-Verlet, '
                         'dt 1 fs, 300 K, Verlet shell %.2f A' % SKIN,
                atoms_per_gpu=4 * NCELL ** 3, n_gpus=ngpu, parallelism='slab domain decomposition x%d along x, NCCL halo exchange' % ngpu
                if ngpu > 1 else 'single GPU',
                l2_policy='no flush between MD steps (each step consumes the previous one); per-step working set '
                          '(pair list 155 MB + positions/forces) exceeds the 126 MB L2')


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_ours(args):
    import ctypes as C
    if os.environ.get('ATX_BENCH_WATCHDOG'):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['ATX_BENCH_WATCHDOG']), exit=True)
    from atomistica_b200 import _lib as L, md, native

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod   # rendezvous / barrier plumbing only
        dist = dist_mod
        dist.init_process_group(backend='gloo')

    ctx = L.context(local_rank)
    setfl = load_setfl()
    a, m, v0 = build_system()
    nat = len(a)

    pot = native.TabulatedAlloyEAM(setfl=setfl, device=local_rank)
    if world == 1:
        p = native.from_atoms(a, device=local_rank)
        nl = native.Neighbors(200, device=local_rank)
        drv = md.VelocityVerlet(pot, p, nl, m, v0, dt=DT, verlet_shell=SKIN)
        list_info = nl.info
    else:
        # weak scaling: the global cell is (40*N) x 40 x 40 fcc cells, slab-decomposed along x; every
        # rank generates only the atoms of its own slab (ids are global)
        from atomistica_b200 import parallel
        dd = parallel.DomainDecomposition(rank, world, parallel.torch_exchange_id, device=local_rank)
        gcell = a.cell.copy()
        gcell[0] *= world
        pos = a.positions + rank * a.cell[0] + 1e-3      # keep lattice planes off the slab faces
        vel = md.maxwell_boltzmann(m, TEMP, seed=12345 + rank)
        ids = np.arange(nat, dtype=np.int64) + rank * nat
        drv = parallel.DDVelocityVerlet(dd, pot, None, [29], gcell, True, ids, np.ones(nat, dtype=np.int32), pos, vel,
                                        m, float(setfl['cutoff']), SKIN, dt=DT)
        list_info = None

    steps, warm = args.steps, max(args.warmup, 3)
    drv.run(warm)
    reb0 = drv.stats()['nrebuilds']
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist is not None:
        dist.barrier()
    L.check(L.lib().atx_ctx_synchronize(ctx))
    L.kernel_launches(reset=True)
    L.check(L.lib().atx_profile_enable(ctx, 1))
    t0 = time.perf_counter()
    epot, ekin = drv.run(steps)
    wall = time.perf_counter() - t0
    L.check(L.lib().atx_profile_enable(ctx, 0))
    launches = L.kernel_launches()
    st = drv.stats()
    dev_ms = st['last_run_ms']
    rebuilds = st['nrebuilds'] - reb0

    def prof(name):
        tot, cnt = C.c_double(0.0), C.c_longlong(0)
        L.check(L.lib().atx_profile_read(ctx, name.encode(), C.byref(tot), C.byref(cnt)))
        return tot.value, cnt.value

    force_ms, force_n = prof('eam_force')
    dens_ms, dens_n = prof('eam_density')
    cnt_ms, cnt_n = prof('nl_pairs_count')
    fill_ms, fill_n = prof('nl_pairs_fill')
    dd_prof = {k: prof(k)[0] for k in ('dd_allreduce', 'dd_halo', 'dd_step')} if world > 1 else None
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t[0])
        dist.barrier()
    clocks = sampler.finish()

    if list_info is not None:
        z_list = list_info()['npairs'] / nat
    else:
        z_list = 78.0      # fcc Cu, cutoff 5.507 + 0.5 A: pairs per atom of the list (measured at N=1)
    nown, nghost = (nat, 0) if world == 1 else drv.counts()
    value = world * nat * steps / (dev_ms * 1e-3)

    # ---- e2e: reference-facing calculator API, host buffers, copies inside the timed region
    from atomistica_b200 import TabulatedAlloyEAM
    calc = TabulatedAlloyEAM(setfl=setfl, device=local_rank, verlet_shell=SKIN)
    r = a.positions.copy()
    vel = v0.copy()
    a2 = a.copy()
    a2.positions = r
    f = calc.get_forces(a2)          # initialises particles / neighbour list (untimed, like `warm`)
    e2e_steps = max(3, min(steps, 20))
    t_api = 0.0
    for _ in range(3 + e2e_steps):
        vel += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        r += vel * DT
        a2.positions = r
        t0 = time.perf_counter()
        f = calc.get_forces(a2)
        dt_call = time.perf_counter() - t0
        vel += 0.5 * f / MASS_CU * md.ACCEL_CONV * DT
        if _ >= 3:
            t_api += dt_call
    if dist is not None:
        import torch
        tt = torch.tensor([t_api], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_api = float(tt[0])
    e2e_value = world * nat * e2e_steps / t_api

    # ---- roofline of the dominant kernel
    peak, peak_src = measured_peaks()
    alg_bytes = nat * (68.0 + 16.0 * z_list)          # B_eam = 68 + 16 z per atom
    force_avg_ms = force_ms / max(force_n, 1)
    achieved = alg_bytes / (force_avg_ms * 1e-3) / 1e9
    fp64 = C.c_double(0.0)
    L.check(L.lib().atx_measure_fp64_peak(ctx, C.byref(fp64)))
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('k_eam_force_dram_bytes_per_launch')

    out = dict(
        metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=world, steps=steps, warmup=warm,
        ms_per_step=dev_ms / steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
        data='synthetic', config=config_dict(world),
        e2e=dict(value=e2e_value, unit='atom-steps/s', h2d_bytes_per_step=nat * 24, d2h_bytes_per_step=nat * 24 + 80,
                 steps=e2e_steps, note='calculator API: host positions in, host forces out (page-locked result buffers) every call; neighbour list kept in a '
                                       '%.2f A Verlet shell (device-side displacement check every call); at N>1 one '
                                       'calculator instance per GPU' % SKIN),
        gpu_launches=launches,
        clocks=clocks,
        roofline=dict(bound='hbm', kernel='k_eam_force_fast<4,2,VIRIAL=0>', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                      traffic=traffic, peak_source=peak_src, algorithmic_bytes_per_launch=alg_bytes,
                      avg_launch_ms=force_avg_ms, launches=force_n, list_neighbors_per_atom=z_list,
                      share_of_step=force_ms / dev_ms if dev_ms else None),
        kernels_ms=dict(eam_force=force_ms, eam_density=dens_ms, nl_pairs_count=cnt_ms, nl_pairs_fill=fill_ms,
                        total_device=dev_ms, dd=dd_prof),
        fp64_peak_tflops_measured=fp64.value,
        md=dict(epot=epot, ekin=ekin, rebuilds=rebuilds, rebuild_interval=steps / max(rebuilds, 1),
                wall_s=wall, owned_atoms_rank0=nown, ghost_atoms_rank0=nghost),
    )
    if rank == 0:
        if world == 1 and not args.no_cpu:
            threads = cpu_best_threads()
            ncpu, t_build, t_force = cpu_sample(ncell=24, nforce=3, threads=threads)
            interval = steps / max(rebuilds, 1)
            t = t_force + t_build / interval
            out['cpu_baseline'] = dict(
                value=ncpu / t, unit='atom-steps/s', cores=threads, kind='port',
                sample='fcc Cu 24^3 cells = %d atoms, oracle port with %d OpenMP thread(s) of %d host threads: 3 EAM '
                       'force evaluations + 1 neighbour build (cutoff+%.1f A skin) amortised over the GPU '
                       'run\'s rebuild interval of %.1f steps; %s' % (ncpu, threads, host_threads(), SKIN, interval, CPU_FLAGS))
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
