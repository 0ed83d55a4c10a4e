#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.
   python tools/ncu_summary.py launches <launches.csv> > profiles/xxx_launches.txt
   python tools/ncu_summary.py full <report.ncu-rep> ... > profiles/xxx_full.txt"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    data = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
    tot = sum(t for _, t in data)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in data:
        k = n.split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:80]
        agg[k][0] += 1
        agg[k][1] += t
    print('# %d launches, %.1f us total (gpu__time_duration.sum, ns in the csv)' % (len(data), tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-82s %5d %11.1f us %5.1f%%' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))


def full(paths):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, u = rows[0], rows[1]
        for v in rows[2:]:
            print('== %s :: %s' % (p.split('/')[-1], v[h.index('Kernel Name')][:90]))
            for k in KEYS:
                if k in h:
                    print('   %-95s %-14s %s' % (k, u[h.index(k)], v[h.index(k)]))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
