#!/usr/bin/env python
"""Debug: per-k-block clock64 timeline of CTA 0 of the tcgen05 GEMM (library built with -DTCG_TIMELINE).
rows: 0 loop top, 1 after cp.async wait, 2 after empty wait, 3 after convert+arrive (producer thread 0);
      4 MMA loop top, 5 after full wait, 6 after MMA issue+commit (MMA thread)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from speedplusbaseline_b200 import _lib as L
from kutil import BnF, sp, vt_bnact
M, N, K = (int(v) for v in sys.argv[1].split(','))
x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); y = torch.empty(M, N, device='cuda')
sc, sh = torch.rand(K, device='cuda') + .5, torch.randn(K, device='cuda') * .1
bnf = BnF(N); xvt = vt_bnact(x, sc, sh, L.ACT_RELU6)
for _ in range(3):
    L.call('b200sp_pw_fwd', C.byref(xvt), w.data_ptr(), None, 0, y.data_ptr(), bnf.ref(), M, N, K, L.F32, sp())
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 256))()
L.lib.b200sp_tcg_timeline.argtypes = [C.c_void_p]; L.lib.b200sp_tcg_timeline(buf)
t = [[buf[r * 256 + i] for i in range(256)] for r in range(8)]
t0 = t[0][0]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
print('kb   top  cpwait  empty  conv | mtop  full  issue   (cycles since start; deltas per stage)')
for i in range(n):
    print('%3d %6d %6d %6d %6d | %6d %6d %6d   d_cp %5d d_empty %5d d_conv %5d | d_full %5d d_issue %5d' % (
        i, t[0][i] - t0, t[1][i] - t0, t[2][i] - t0, t[3][i] - t0, t[4][i] - t0, t[5][i] - t0, t[6][i] - t0,
        t[1][i] - t[0][i], t[2][i] - t[1][i], t[3][i] - t[2][i], t[5][i] - t[4][i], t[6][i] - t[5][i]))
