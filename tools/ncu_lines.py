#!/usr/bin/env python
"""Per-source-line summary of an ncu report captured with --import-source on:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv ; python tools/ncu_lines.py X.csv [N]
prints warp-instructions, stall samples and the two dominant stall reasons of the N hottest source lines."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    cur, h, agg = None, None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            h = r
            continue
        if not r or h is None or len(r) < len(h) or not r[0].isdigit():
            continue
        try:
            s, ie = int(r[6]), int(r[7])
        except ValueError:
            continue
        st = {h[i]: int(r[i]) for i in range(32, 49) if r[i].isdigit()}
        agg[(cur, int(r[0]))] = (s, ie, r[1], st)
    tot_s, tot_i = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
    print('# %d stall samples, %d warp-instructions' % (tot_s, tot_i))
    allst = collections.Counter()
    for v in agg.values():
        allst.update(v[3])
    print('# stall mix: ' + ', '.join('%s %.0f%%' % (k.replace('stall_', ''), 100.0 * n / tot_s) for k, n in allst.most_common(6)))
    for key in (1, 0):
        print('# --- top lines by ' + ('instructions' if key else 'stall samples'))
        for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top_n]:
            top = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
            print('%-16s %4d  inst %5.1f%%  samples %5.1f%%  %-34s | %s' % (
                f[:16], l, 100.0 * v[1] / tot_i, 100.0 * v[0] / tot_s,
                ' '.join('%s=%d' % (k.replace('stall_', ''), n) for k, n in top), v[2].strip()[:84]))


if __name__ == '__main__':
    main()
