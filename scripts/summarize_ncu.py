#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries in profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/r01_launches_final.csv profiles/r01_launches_final.csv "<command>"
    python scripts/summarize_ncu.py full gpurun_out/r01_eam_final.ncu-rep profiles/r01_ncu_eam_final.csv
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    'Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__inst_executed.sum',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
    'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
    'launch__occupancy_limit_warps', 'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_bytes.sum',
    'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_op_global_red.sum',
]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src, errors='replace')) if r and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = r[4].split('(')[0]
        val = float(r[-1].replace(',', ''))
        unit = r[-2]
        ns = val * dict(ns=1.0, us=1e3, usecond=1e3, nsecond=1.0, ms=1e6, msecond=1e6).get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    out = ['# ' + cmd, '# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes',
           'kernel,launches,total_ms,avg_us,share_pct']
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append('"%s",%d,%.3f,%.1f,%.1f' % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, v[1] / tot * 100))
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:14]))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    want = [m for m in METRICS if m in h]
    idx = [h.index(m) for m in want]
    out = ['# ncu --set full --clock-control none --import-source on; source report: %s' % os.path.basename(src),
           '# units: ' + ','.join(rr[1][i] for i in idx), ','.join(want)]
    for row in rr[2:]:
        out.append(','.join('"%s"' % row[i].split('(')[0] if k == 0 else row[i] for k, i in enumerate(idx)))
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))
    return rr, h


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
    else:
        rr, h = full(sys.argv[2], sys.argv[3])
        # DRAM traffic of the dominant kernel for bench.py's roofline.traffic
        ik, ir, iw = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
        scale = dict(Mbyte=1e6, Kbyte=1e3, Gbyte=1e9, byte=1.0)
        vals = [float(r[ir]) * scale[rr[1][ir]] + float(r[iw]) * scale[rr[1][iw]] for r in rr[2:]
                if r[ik].startswith('void k_eam_force_fast')]
        if vals:
            tp = os.path.join(ROOT, 'profiles', 'traffic.json')
            json.dump(dict(k_eam_force_dram_bytes_per_launch=sum(vals) / len(vals),
                           source=os.path.basename(sys.argv[3]), launches_averaged=len(vals)), open(tp, 'w'), indent=1)
            print('traffic.json:', sum(vals) / len(vals))
