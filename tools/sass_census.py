#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that identify the execution path (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync, LDGSTS = cp.async.
    python tools/sass_census.py build/obj/*.o        (cuobjdump -sass on the objects _build.py leaves behind; no GPU needed)"""
import collections
import re
import subprocess
import sys

KEYS = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'SYNCS', 'ATOMS.CAST', 'RED.E', 'FFMA', 'DFMA')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return [re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', o) for o in out]


def main():
    for obj in sys.argv[1:]:
        txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        funcs = re.split(r'\n\s*Function : ', txt)[1:]
        names = demangle([f.split('\n', 1)[0].strip() for f in funcs])
        agg = collections.OrderedDict()
        for name, f in zip(names, funcs):
            base = re.sub(r'\(.*$', '', name)
            base = re.sub(r'^void ', '', base)
            c = agg.setdefault(base.split('<')[0], collections.Counter())
            c['instantiations'] += 1
            for l in f.split('\n'):
                m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
                if m:
                    op = m.group(1)
                    for k in KEYS:
                        if op.startswith(k):
                            c[k] += 1
        print('== ' + obj.split('/')[-1])
        for k, c in agg.items():
            n = c.pop('instantiations')
            print('  %-28s x%-3d %s' % (k, n, '  '.join('%s %d' % (kk, c[kk]) for kk in KEYS if c[kk])))


if __name__ == '__main__':
    main()
