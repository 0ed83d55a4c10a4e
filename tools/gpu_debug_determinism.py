#!/usr/bin/env python
"""Bitwise run-to-run determinism of the 1x1-convolution outputs (y of pw_fwd, dx of pw_dgrad have no atomics: they must repeat
exactly) on the KRN shapes at batch 4; B200SP_TCG2_LEAN=off etc. select the kernel variant."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from speedplusbaseline_b200 import _lib as L
from kutil import BnB, BnF, sp, vt_bnact, vt_dy
L.ensure_workspace('cuda')
dev = 'cuda'
SH = [(50176, 16, 32), (50176, 96, 16), (12544, 24, 96), (12544, 144, 24), (12544, 24, 144), (3136, 32, 144), (3136, 192, 32), (3136, 32, 192),
      (784, 64, 192), (784, 384, 64), (784, 64, 384), (784, 96, 384), (784, 576, 96), (784, 96, 576), (196, 160, 576), (196, 960, 160),
      (196, 160, 960), (196, 320, 960), (196, 1024, 320), (196, 1024, 1024), (196, 1024, 1280)]
bad = 0
for M, N, K in SH:
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    g = torch.randn(M, N, device=dev); yq = torch.randn(M, N, device=dev)
    sck, shk = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
    cA, cB, cC = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev) * 0.1, torch.randn(N, device=dev) * 0.1
    xvt = vt_bnact(x, sck, shk, L.ACT_RELU6); dyvt = vt_dy(g, yq, cA, cB, cC)
    outs = {'fwd': [], 'dgrad': []}
    for rep in range(12):
        y = torch.full((M, N), float('nan'), device=dev); gx = torch.full((M, K), float('nan'), device=dev)
        bnf = BnF(N); bnb = BnB(x, sck, shk, torch.zeros(K, device=dev), torch.ones(K, device=dev), L.ACT_RELU6)
        L.call('b200sp_pw_fwd', C.byref(xvt), w.data_ptr(), None, 0, y.data_ptr(), bnf.ref(), M, N, K, L.F32, sp())
        L.call('b200sp_pw_dgrad', C.byref(dyvt), w.data_ptr(), None, 1.0, gx.data_ptr(), bnb.ref(), M, N, K, L.F32, sp())
        torch.cuda.synchronize()
        outs['fwd'].append(y); outs['dgrad'].append(gx)
    for op, lst in outs.items():
        nd = sum(int(not torch.equal(lst[0], t)) for t in lst[1:])
        nn = int(torch.isnan(lst[0]).sum())
        if nd or nn:
            bad += 1
            d = max(float((lst[0] - t).abs().max()) for t in lst[1:])
            print('%-6s [%d,%d,%d]: %d of 11 repeats differ (max |diff| %.3e), NaN %d' % (op, M, N, K, nd, d, nn))
print('non-deterministic shapes:', bad, {k: v for k, v in os.environ.items() if k.startswith('B200SP_TCG2')})
