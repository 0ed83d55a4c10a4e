#!/usr/bin/env python
"""Micro-benchmark of the 1x1-conv GEMM entry points (pw_fwd / pw_dgrad / pw_wgrad) on the KRN shapes
(SURVEY.md Appendix A.1): CUDA-event time per launch, algorithmic GB/s and TFLOP/s.
    python tools/gemm_bench.py [--shapes M,N,K ...] [--reps 20] [--ops fwd,dgrad,wgrad]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from speedplusbaseline_b200 import _lib as L          # noqa: E402
from kutil import BnB, BnF, sp, vt_bnact, vt_dy, vt_plain   # noqa: E402

KRN = [(602112, 16, 32), (602112, 96, 16), (150528, 24, 96), (150528, 144, 24), (150528, 24, 144), (37632, 32, 144),
       (37632, 192, 32), (37632, 32, 192), (9408, 64, 192), (9408, 384, 64), (9408, 64, 384), (9408, 96, 384),
       (9408, 576, 96), (9408, 96, 576), (2352, 160, 576), (2352, 960, 160), (2352, 160, 960), (2352, 320, 960),
       (2352, 1024, 320), (2352, 1024, 1024), (2352, 1024, 1280)]


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def timeit_graph(fn, reps):
    """`reps` back-to-back calls captured in ONE CUDA graph (what the train step does): no host launch cost between the kernels,
    programmatic dependent launch active, operands L2-warm.  Returns us per call."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shapes', nargs='*', default=None)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--ops', default='fwd,dgrad,wgrad')
    ap.add_argument('--graph', action='store_true', help='time the calls inside one CUDA graph (no host launch gaps, L2-warm)')
    args = ap.parse_args()
    shapes = KRN if not args.shapes else [tuple(int(v) for v in s.split(',')) for s in args.shapes]
    ops = args.ops.split(',')
    dev = 'cuda'
    if os.environ.get('B200SP_WS', '1') != '0':       # presplit route of the <= 4096-row layers (B200SP_WS=0: general kernel only)
        L.ensure_workspace(dev)
    print('%-22s %-6s %9s %9s %9s' % ('shape [M,N,K]', 'op', 'us', 'GB/s', 'TFLOP/s'))
    for M, N, K in shapes:
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
        flops = 2.0 * M * N * K
        runs = {
            'fwd': (lambda: L.call('b200sp_pw_fwd', C.byref(xvt), w.data_ptr(), None, 0, y.data_ptr(), bnf.ref(), M, N, K, L.F32, sp()),
                    (M * K + M * N + N * K) * 4),
            'dgrad': (lambda: L.call('b200sp_pw_dgrad', C.byref(dyvt), w.data_ptr(), None, 1.0, gx.data_ptr(), bnb.ref(), M, N, K, L.F32, sp()),
                      (2 * M * N + 2 * M * K + N * K) * 4),
            'wgrad': (lambda: L.call('b200sp_pw_wgrad', C.byref(dyvt), C.byref(xvt), dw.data_ptr(), None, M, N, K, L.F32, sp()),
                      (2 * M * N + M * K + 2 * N * K) * 4),
        }
        for op in ops:
            fn, nbytes = runs[op]
            us = timeit_graph(fn, max(args.reps, 20)) if args.graph else timeit(fn, args.reps)
            print('%-22s %-6s %9.1f %9.1f %9.2f' % ('[%d,%d,%d]' % (M, N, K), op, us, nbytes / us / 1e3, flops / us / 1e6), flush=True)


if __name__ == '__main__':
    main()
