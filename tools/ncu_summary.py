#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): key raw metrics, instruction mix and stall reasons from the source page."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in keys:
        print("%-70s %-14s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop, samples, stall = collections.Counter(), collections.Counter(), collections.Counter()
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = r[ix['Source']].strip()
    n = int(r[ix['Instructions Executed']])
    op = (s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    byop[op] += n
    samples[op] += int(r[ix['# Samples']])
    tot += n
    for c in stall_cols:
        stall[c] += int(r[ix[c]])
print("total warp instructions", tot)
for op, n in byop.most_common(16):
    print("  %-10s %11d %5.1f%%  samples %d" % (op, n, 100.0 * n / tot, samples[op]))
ts = sum(stall.values())
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / ts) for k, v in stall.most_common(9)))
