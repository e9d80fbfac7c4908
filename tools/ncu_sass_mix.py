#!/usr/bin/env python
"""Opcode mix and most-stalled SASS lines of the first kernel in an .ncu-rep.
usage: python tools/ncu_sass_mix.py rep.ncu-rep units_per_launch [n_lines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = None
ops, samp = collections.Counter(), collections.Counter()
tot_i = tot_s = 0
lines = []
for r in rows:
    if r and r[0] == 'Address':
        if h is not None:
            break  # first kernel only
        h = r
        continue
    if h and len(r) == len(h):
        d = dict(zip(h, r))
        src = d.get('Source', '')
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
        op = m.group(2).split('.')[0] if m else '?'
        ie = int(float(d.get('Instructions Executed', '0') or 0))
        s = int(float(d.get('# Samples', '0') or 0))
        ops[op] += ie
        samp[op] += s
        tot_i += ie
        tot_s += s
        lines.append((s, ie, src.strip()[:90]))
print('total warp instr', tot_i, 'per unit', tot_i / units, 'samples', tot_s)
for op, c in ops.most_common(30):
    print(f'{op:12s} inst/unit {c / units:7.1f}  samples {100 * samp[op] / max(1, tot_s):5.1f}%')
print('--- most stalled SASS lines')
for s, ie, src in sorted(lines, reverse=True)[:nl]:
    print(f'{100 * s / max(1, tot_s):5.1f}% {ie / units:6.2f}/unit  {src}')
