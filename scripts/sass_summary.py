"""Static SASS summary of the hot kernels in libatomistica_b200.so (cuobjdump, no GPU needed): registers, shared
memory and counts of the mnemonics that matter for the design claims -- FP64 arithmetic, FP64 atomics
(RED/ATOM .F64), 256-bit loads, shuffles, votes.  Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'atomistica_b200', 'libatomistica_b200.so')
KEEP = ('k_eam_force_fast', 'k_eam_density_fast', 'k_bop_center', 'k_bop_gather', 'k_bopscr_center', 'k_bopscr_bonds',
        'k_rebo2_force_bond', 'k_rebo2_bonds', 'k_rbs_force', 'k_pairs_coop', 'k_pairs_f32', 'k_md_kickdrift',
        'k_dd_kickdrift', 'k_dd_pack_p2p', 'k_dd_wait', 'k_an_histogram')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


res = subprocess.run(['cuobjdump', '--dump-resource-usage', SO], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r'\s*Function (\S+):', line)
    if m:
        cur = m.group(1)
    elif cur and 'REG:' in line:
        usage[cur] = dict(re.findall(r'(REG|STACK|SHARED):(\d+)', line))
        cur = None
sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c['total'] += 1
        for key, pat in (('DFMA', r'^DFMA'), ('DADD', r'^DADD'), ('DMUL', r'^DMUL'), ('MUFU', r'^MUFU'),
                         ('RED/ATOM.F64', r'^(REDG?|ATOMG?)\..*F64'), ('LDG.256', r'^LDG\..*256'),
                         ('LDG', r'^LDG'), ('LDS', r'^LDS'), ('STS', r'^STS'), ('SHFL', r'^SHFL'), ('VOTE', r'^VOTE'),
                         ('FFMA', r'^FFMA')):
            if re.match(pat, op):
                c[key] += 1
names = [k for k in counts if any(s in k for s in KEEP)]
dm = demangle(names)
cols = ['total', 'DFMA', 'DADD', 'DMUL', 'MUFU', 'FFMA', 'RED/ATOM.F64', 'LDG', 'LDG.256', 'LDS', 'STS', 'SHFL', 'VOTE']
print('# cuobjdump -sass / --dump-resource-usage of atomistica_b200/libatomistica_b200.so (static instruction counts per '
      'kernel, not executed counts)')
print('kernel,registers,stack,shared_static,' + ','.join(cols))
for k in sorted(names, key=lambda x: dm[x]):
    short = re.sub(r'\(.*', '', dm[k]).replace('void ', '')
    u = usage.get(k, {})
    print('"%s",%s,%s,%s,' % (short, u.get('REG', ''), u.get('STACK', ''), u.get('SHARED', '')) +
          ','.join(str(counts[k][c]) for c in cols))
