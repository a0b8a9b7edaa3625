#!/usr/bin/env python
"""Read an .ncu-rep here (no GPU): opcode mix and execution-count segments of the SASS of the
first kernel in the report.  usage: ncu_segments.py report.ncu-rep [min_share_percent]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h) - 2]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


def opcode(r):
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
    return m.group(2).split('.')[0] if m else '?'


ti = sum(f(r, 'Instructions Executed') for r in data)
ts = sum(f(r, '# Samples') for r in data)
tt = sum(f(r, 'Thread Instructions Executed') for r in data)
print(f'{len(data)} SASS instructions ({len(data) * 16 / 1024:.1f} KB), {ti:.4g} warp instructions '
      f'executed, {tt / ti:.2f} threads per instruction')
byop, bysm, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    op = opcode(r)
    byop[op] += f(r, 'Instructions Executed')
    bysm[op] += f(r, '# Samples')
    thr[op] += f(r, 'Thread Instructions Executed')
print('opcode      inst%  samples%  threads/inst')
for op, c in byop.most_common(18):
    print(f'{op:10s} {c / ti * 100:6.1f} {bysm[op] / ts * 100:8.1f} {thr[op] / max(c, 1):10.1f}')
seg, cur = [], None
for i, r in enumerate(data):
    c = f(r, 'Instructions Executed')
    if cur is None or abs(c - cur['c']) > 0.02 * max(c, cur['c'], 1):
        cur = {'c': c, 'start': i, 'n': 0, 'thr': 0.0, 'smp': 0.0, 'ops': collections.Counter()}
        seg.append(cur)
    cur['n'] += 1
    cur['thr'] += f(r, 'Avg. Threads Executed')
    cur['smp'] += f(r, '# Samples')
    cur['ops'][opcode(r)] += 1
print('segments (consecutive SASS with the same execution count):')
for s in seg:
    share = s['c'] * s['n'] / ti * 100
    if share > min_share:
        print(f"sass[{s['start']:4d}+{s['n']:4d}] exec {s['c']:.3g} share {share:5.1f}% samples "
              f"{s['smp'] / ts * 100:5.1f}% threads {s['thr'] / s['n']:5.1f} "
              f"{dict(s['ops'].most_common(4))}")
