#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed warp-instructions and
stall samples per SASS opcode, plus the hottest instructions."""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
thr = collections.Counter()
smp = collections.Counter()
tot = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ci['Source']].strip()
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
    if not m:
        continue
    op = m.group(1)
    ex = int(r[ci['Instructions Executed']] or 0)
    te = int(r[ci['Thread Instructions Executed']] or 0)
    sa = int(r[ci['# Samples']] or 0)
    ops[op] += ex
    thr[op] += te
    smp[op] += sa
    tot += ex
    recs.append((sa, ex, src))
print(f'total warp-instructions executed: {tot:.4g}   thread-instr: {sum(thr.values()):.4g}')
if steps:
    print(f'thread-instr per attempted step: {sum(thr.values()) / steps:.1f}')
tsmp = sum(smp.values())
print(f'{"op":12s} {"warp-inst":>12s} {"%":>6s} {"thr/inst":>8s} {"samples%":>8s}')
for op, ex in ops.most_common(28):
    print(f'{op:12s} {ex:12.4g} {100 * ex / tot:6.2f} {thr[op] / max(ex, 1):8.1f} {100 * smp[op] / max(tsmp, 1):8.2f}')
print('\nhottest instructions by stall samples:')
for sa, ex, src in sorted(recs, reverse=True)[:25]:
    print(f'{100 * sa / tsmp:6.2f}%  ex={ex:10d}  {src[:90]}')
