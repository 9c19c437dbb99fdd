#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into per-kernel shares.
    python tools/launch_summary.py gpurun_out/r01b_launches.csv [tail_fraction]
The tail of the list is the graph replays of the refine iteration (the reconstruction and the capture come first)."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    tail = float(sys.argv[2]) if len(sys.argv) > 2 else 0.45
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    for r in rd:
        if len(r) > iv and r[im] == 'gpu__time_duration.sum':
            rows.append((r[ik], float(r[iv].replace(',', ''))))
    n = len(rows)
    sel = rows[int(n * (1 - tail)):]
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in sel:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    ours = sum(v for k, (_, v) in agg.items() if 'lf::' in k or 'tc::' in k or 'dz::' in k or k.startswith('lf') or 'conv_tc' in k)
    print(f"{n} launches in the whole command; table = last {tail:.0%} of launches (graph replays of the refine iteration), "
          f"durations are cold-cache/serialised (ns): compare SHARES")
    print(f"share of lfb200 kernels: {100 * ours / tot:.1f}%   (torch glue {100 * (1 - ours / tot):.1f}%)\n")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{100 * v / tot:6.2f}%  n={c:4d}  avg={v / c / 1e3:9.1f} us  {k[:90]}")


if __name__ == '__main__':
    main()
