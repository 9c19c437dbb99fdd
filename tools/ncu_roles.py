#!/usr/bin/env python
"""Bucket ncu source-page samples of conv_tc_kernel by warp role (uses the mbarrier/tcgen05 markers)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    try: data.append((int(r[isamp]), int(r[iex]), r[isrc]))
    except Exception: pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
width = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for b in range(0, len(data), width):
    seg = data[b:b + width]
    s = sum(d[0] for d in seg); e = sum(d[1] for d in seg)
    marks = collections.Counter()
    for _, _, src in seg:
        for m in ('UTCHMMA', 'TRYWAIT', 'LDTM', 'LDG', 'STS', 'STG', 'UTCBAR', 'F2F', 'BAR.SYNC', 'ARRIVE'):
            if m in src: marks[m] += 1
    print(f'{b:5d} samples={s:6d} ({100*s/max(tot,1):4.1f}%) exec={e:10d} {dict(marks)}')
top = sorted(enumerate(data), key=lambda t: -t[1][0])[:25]
for i, (s, e, src) in top:
    print(f'{i:5d} {s:6d} {100*s/tot:4.1f}% ex={e:9d} {src[:100]}')
