#!/usr/bin/env python
"""Summarise the CSV pages written by tools/ncu_capture.sh:
opcode mix / stall samples of the hot loop and headline raw metrics."""
import csv
import gzip
import sys
from collections import Counter


def main(tag, hot_count, full=False):
    raw = list(csv.reader(open('gpurun_out/prof_%s_raw.csv' % tag)))
    hdr, vals = raw[0], raw[2]
    keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max',
            'smsp__inst_executed.sum', 'launch__registers_per_thread',
            'launch__grid_size', 'launch__block_size',
            'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'smsp__average_warp_latency_per_inst_issued.ratio',
            'sm__inst_executed_pipe_fp64.sum',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
    for k in keys:
        if k in hdr:
            print('%-60s %s %s' % (k, vals[hdr.index(k)], raw[1][hdr.index(k)]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') \
                and 'not_issued' not in h:
            try:
                if float(vals[i]) > 0.05:
                    print('  stall %-28s %s' % (h.replace(
                        'smsp__average_warps_issue_stalled_', '').replace(
                            '_per_issue_active.ratio', ''), vals[i]))
            except ValueError:
                pass
    rows = list(csv.reader(gzip.open(
        'gpurun_out/prof_%s_source.csv.gz' % tag, 'rt')))
    hdr, data = rows[1], rows[2:]
    ia, ie, isamp = hdr.index('Source'), hdr.index(
        'Instructions Executed'), hdr.index('# Samples')
    cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
    ci = [hdr.index(c) for c in cols]
    ops, samp, hot, tot_s, lines = Counter(), Counter(), 0, 0, []
    for r in data:
        try:
            n, s = int(r[ie]), int(r[isamp])
        except ValueError:
            continue
        tot_s += s
        if n >= 0.7 * hot_count:
            toks = r[ia].split()
            op = toks[1] if toks[0].startswith('@') else toks[0]
            ops[op.split('.')[0]] += n / hot_count
            samp[op.split('.')[0]] += s
            hot += n / hot_count
            st = ' '.join('%s=%s' % (c[6:], r[i]) for c, i in zip(cols, ci)
                          if r[i] not in ('0', ''))
            lines.append((s, n, r[ia], st))
    print('hot-loop instructions per step: %.0f ; total samples %d ; SASS '
          'lines %d' % (hot, tot_s, len(data)))
    for k, v in ops.most_common(18):
        print('  %-12s %7.1f/step  samples %d' % (k, v, samp[k]))
    if full:
        for i, (s, n, a, st) in enumerate(lines):
            print('%4d %3d %-58s %s' % (i, s, a[:58], st))
    else:
        print('top sampled:')
        for s, n, a, st in sorted(lines, reverse=True)[:16]:
            print('  %3d %-58s %s' % (s, a[:58], st))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]), len(sys.argv) > 3)
