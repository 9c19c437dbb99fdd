#!/usr/bin/env python
"""Summarise one .ncu-rep: key raw metrics + the instructions with the most stall samples (dev tool)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:100])
    for k in KEYS:
        if k in hdr:
            print(f'  {k} = {r[hdr.index(k)]} {unit[hdr.index(k)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if r and r[0] == 'Address':
        h = r; data = []; continue
    if h and len(r) == len(h):
        data.append(r)
if h:
    ix = {k: i for i, k in enumerate(h)}
    st = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    tot = sum(int(r[ix['# Samples']] or 0) for r in data)
    agg = sorted(((sum(int(r[ix[s]] or 0) for r in data), s) for s in st), reverse=True)[:7]
    print('samples', tot, 'instr', sum(int(r[ix['Instructions Executed']] or 0) for r in data), agg)
    for k in ('L1 Wavefronts Shared', 'L1 Tag Requests Global'):
        if k in ix: print(k, sum(float(r[ix[k]] or 0) for r in data))
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
        m = sorted(((int(r[ix[s]] or 0), s) for s in st), reverse=True)[:2]
        print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']][:95], m)
