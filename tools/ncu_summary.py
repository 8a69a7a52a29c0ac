#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + hottest SASS lines (by samples / inst)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__lsuin_requests.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('===', r[hdr.index('Kernel Name')][:80])
    for k in keys:
        if k in hdr:
            print(f'  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[h]
ii, si, st = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
stall_cols = [j for j, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for r in rows[h + 1:]:
    if len(r) > ii and r[ii].isdigit():
        data.append((int(r[ii]), int(r[st]), r[si].strip(), [int(r[j] or 0) for j in stall_cols]))
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f'total warp-inst {tot}  samples {ts}  sass lines {len(data)}')
agg = [0] * len(stall_cols)
for d in data:
    for j, v in enumerate(d[3]): agg[j] += v
print('stall mix:', ', '.join(f'{hdr[c][6:]}={v / max(1, sum(agg)) * 100:.1f}%' for c, v in sorted(zip(stall_cols, agg), key=lambda t: -t[1])[:8]))
# opcode mix
mix = {}
for v, s, srcl, _ in data:
    op = srcl.split()[1] if srcl.startswith('@') else srcl.split()[0]
    op = op.split('.')[0]
    mix[op] = mix.get(op, 0) + v
print('opcode mix:', ', '.join(f'{k}={v / tot * 100:.1f}%' for k, v in sorted(mix.items(), key=lambda t: -t[1])[:16]))
print('hottest by samples:')
for k, (v, s, srcl, stl) in sorted(enumerate(data), key=lambda t: -t[1][1])[:top]:
    why = max(zip(stl, stall_cols))[1]
    print(f'  {k:5d} inst {v / tot * 100:5.2f}%  smp {s / ts * 100:5.2f}%  {hdr[why][6:]:14s} {srcl[:80]}')
