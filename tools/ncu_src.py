#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: samples / instructions per SASS region and the top stalls."""
import csv, sys, collections
def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if 'Source' in r][0]
    h = rows[hdr]
    si, sa, ie = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_')]
    data = []
    seen0 = None
    for r in rows[hdr + 1:]:
        if len(r) <= ie: continue
        if seen0 is None: seen0 = r[0]
        elif r[0] == seen0: break          # listing repeats
        try: data.append((len(data), int(r[sa]), int(r[ie]), r[si].strip(), r))
        except ValueError: pass
    return h, data, stall_cols
h, data, stall_cols = load(sys.argv[1])
tot = sum(d[1] for d in data); toti = sum(d[2] for d in data)
print('sass', len(data), 'samples', tot, 'warp-instr', toti)
marks = {d[0]: d[3] for d in data if any(k in d[3] for k in ('LDTM', 'UTCHMMA', 'UTCBAR', 'LDGSTS', 'SYNCS', 'BAR.SYNC', 'NANOSLEEP', 'EXIT', 'UTCATOMSWS'))}
step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reg = collections.OrderedDict()
for n, s, i, src, r in data:
    k = n // step
    reg.setdefault(k, [0, 0]); reg[k][0] += s; reg[k][1] += i
for k, v in reg.items():
    if v[0] or v[1]:
        ms = sorted(set(m.split()[0] if not m.startswith('@') else m.split()[1] for n, m in marks.items() if n // step == k))
        print('%5d  samples %5d (%4.1f%%)  instr %9d (%4.1f%%)  %s' % (k * step, v[0], 100 * v[0] / tot, v[1], 100 * v[1] / toti, ' '.join(ms)[:90]))
print('--- top stalls')
for d in sorted(data, key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print('%5d %5.1f%% %9d  #%d %s' % (d[1], 100 * d[1] / tot, d[2], d[0], d[3][:90]))
