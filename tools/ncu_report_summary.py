#!/usr/bin/env python
"""Summarise one .ncu-rep (read with `ncu -i`): headline raw metrics, stall reasons per issue,
opcode mix and the SASS lines with the most stall samples.
Usage: python tools/ncu_report_summary.py report.ncu-rep out_prefix"""
import csv
import subprocess
import sys
from collections import Counter

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
open(out + '_raw.csv', 'w').write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
lines = ["report: %s" % rep, "kernel: %s" % vals[hdr.index('Kernel Name')]]
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in hdr:
        lines.append('%-62s %s %s' % (k, vals[hdr.index(k)], units[hdr.index(k)]))
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try:
            if float(vals[i]) > 0.05:
                lines.append('  stall %-28s %s' % (h.replace('smsp__average_warps_issue_stalled_', '').replace(
                    '_per_issue_active.ratio', ''), vals[i]))
        except ValueError:
            pass
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
h = srows[1]
isrc, isamp, iex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
data = []
for r in srows[2:]:
    try:
        data.append((int(r[isamp]), int(r[iex]), r[isrc].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
ops = Counter()
for s, e, t in data:
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    ops[op.split('.')[0]] += e
lines.append("SASS lines %d, stall samples %d; warp instructions executed by opcode:" % (len(data), tot))
lines.append("  " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
lines.append("lines with the most stall samples:")
for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:14]):
    lines.append("  %5d  %6d samples (%4.1f %%)  executed %9d  %s" % (
        i, data[i][0], 100.0 * data[i][0] / max(tot, 1), data[i][1], data[i][2][:70]))
open(out + '_summary.txt', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
