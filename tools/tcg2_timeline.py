#!/usr/bin/env python
"""Per-k-block clock64 timeline of CTA 0 of the second-generation tcgen05 GEMM (library built with -DTCG_TIMELINE:
B200SP_LIB_SUFFIX=_tl B200SP_NVCC_EXTRA=-DTCG_TIMELINE python -c "from speedplusbaseline_b200 import _build; _build.build()").
    B200SP_LIB_SUFFIX=_tl B200SP_TCG2=1 python tools/tcg2_timeline.py M,N,K fwd|dgrad|wgrad [rows]
rows of the device buffer: 0/1 TMA thread before/after the rawempty wait; 2 converter loop top, 3 after rawfull, 4 after empty,
5 after convert+fence+arrive (converter thread 0); 6 MMA loop top, 7 after full, 8 after issue+commit; 9 epilogue (2i: accumulator
i observed full, 2i+1: tile i stored)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from speedplusbaseline_b200 import _lib as L          # noqa: E402
from kutil import BnB, BnF, sp, vt_bnact, vt_dy   # noqa: E402

M, N, K = (int(v) for v in sys.argv[1].split(','))
op = sys.argv[2] if len(sys.argv) > 2 else 'fwd'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 24
dev = 'cuda'
x = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) / K ** 0.5
y = torch.empty(M, N, device=dev)
g = torch.randn(M, N, device=dev)
gx = torch.empty(M, K, device=dev)
dw = torch.zeros(N, K, device=dev)
sck, shk = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
cA, cB, cC = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev) * 0.1, torch.randn(N, device=dev) * 0.1
bnf = BnF(N)
bnb = BnB(x, sck, shk, torch.zeros(K, device=dev), torch.ones(K, device=dev), L.ACT_RELU6)
xvt = vt_bnact(x, sck, shk, L.ACT_RELU6)
dyvt = vt_dy(g, y, cA, cB, cC)
run = {'fwd': lambda: L.call('b200sp_pw_fwd', C.byref(xvt), w.data_ptr(), None, 0, y.data_ptr(), bnf.ref(), M, N, K, L.F32, sp()),
       'dgrad': lambda: L.call('b200sp_pw_dgrad', C.byref(dyvt), w.data_ptr(), None, 1.0, gx.data_ptr(), bnb.ref(), M, N, K, L.F32, sp()),
       'wgrad': lambda: L.call('b200sp_pw_wgrad', C.byref(dyvt), C.byref(xvt), dw.data_ptr(), None, M, N, K, L.F32, sp())}[op]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    run()
torch.cuda.synchronize()
buf = (C.c_longlong * (14 * 512))()
L.lib.b200sp_tcg2_timeline.argtypes = [C.c_void_p]
L.lib.b200sp_tcg2_timeline(buf)
t = [[buf[r * 512 + i] for i in range(512)] for r in range(14)]
t0 = min(v for v in (t[0][0], t[2][0], t[6][0]) if v > 0)
print('# %s [%d,%d,%d]  CTA 0, cycles since first stamp' % (op, M, N, K))
print(' kb | tma: wait   issue | conv: top  rawfull   empty    done  (d_raw d_empty d_conv) | mma: top    full   issued (d_full d_issue) | period')
prev = None
for i in range(n):
    if t[5][i] == 0:
        break
    per = (t[8][i] - prev) if prev else 0
    prev = t[8][i]
    print('%3d | %9d %7d | %9d %8d %7d %7d  (%5d %6d %6d) | %8d %7d %8d (%6d %7d) | %6d' % (
        i, t[0][i] - t0, t[1][i] - t0, t[2][i] - t0, t[3][i] - t0, t[4][i] - t0, t[5][i] - t0,
        t[3][i] - t[2][i], t[4][i] - t[3][i], t[5][i] - t[4][i],
        t[6][i] - t0, t[7][i] - t0, t[8][i] - t0, t[7][i] - t[6][i], t[8][i] - t[7][i], per))
print('# epilogue tiles (accumulator observed full -> tile stored):')
for i in range(6):
    if t[9][2 * i] == 0:
        break
    print('   tile %d: full at %8d, stored at %8d (%6d cycles)' % (i, t[9][2 * i] - t0, t[9][2 * i + 1] - t0, t[9][2 * i + 1] - t[9][2 * i]))
print('# epilogue warp 0, per column chunk: start -> +tmem ld -> +correction ld/add -> +stage to smem -> +coalesced stores (+stats after)')
for c in range(12):
    b = t[10][5 * c:5 * c + 5]
    if b[4] == 0:
        break
    nxt = t[10][5 * c + 5]
    print('   chunk %2d: start %8d  ld %5d  corr %5d  stage %5d  store %5d   (next chunk starts +%d)' % (
        c, b[0] - t0, b[1] - b[0], b[2] - b[1], b[3] - b[2], b[4] - b[3], (nxt - b[4]) if nxt else 0))
print('# converter thread 0, inside a k-block: empty observed -> +barrier/recheck -> +lds/convert/sts issued -> +proxy fence -> +arrive')
for i in range(min(n, 12)):
    if t[5][i] == 0:
        break
    print('   kb %2d: bar %5d  convert %5d  fence %5d  arrive %5d' % (i, t[11][i] - t[4][i], t[12][i] - t[11][i], t[13][i] - t[12][i], t[5][i] - t[13][i]))
