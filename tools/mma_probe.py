#!/usr/bin/env python
"""tcgen05.mma issue-rate probe (csrc/mma_probe.cu): cycles per MMA for operand-source / layout / width variants.
    python tools/mma_probe.py > profiles/r2_mma_probe.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speedplusbaseline_b200 import _lib as L          # noqa: E402

KIND = {0: 'tf32', 1: 'bf16'}
SRC = {0: 'A smem', 1: 'A tmem'}
LAY = {0: 'SW128', 1: 'none', 2: 'SW64', 3: 'SW32'}


def run(kind, a_src, layout, N, iters, rotate, grid):
    out = torch.zeros(grid, dtype=torch.int64, device='cuda')
    L.call('b200sp_mma_probe', kind, a_src, layout, N, iters, rotate, grid, out.data_ptr(), L.stream_ptr())
    torch.cuda.synchronize()
    c = out.cpu().double()
    return float(c.mean()) / iters, float(c.max()) / iters


def main():
    iters = 2000
    print('# cycles per tcgen05.mma (M=128, K=32 bytes), %d back-to-back MMAs per CTA; mean / max over CTAs' % iters)
    print('%-5s %-7s %-6s %4s %6s %5s %10s %10s %12s' % ('kind', 'A from', 'layout', 'N', 'rotate', 'grid', 'cyc/mma', 'max', 'MAC/clk/SM'))
    for grid in (1, 148):
        for kind in (0, 1):
            for a_src in (0, 1):
                for layout in (0, 1, 2, 3):
                    for N in (32, 64, 128, 256):
                        for rotate in ((0, 1) if (layout == 0 and N == 128) else (1,)):
                            try:
                                m, mx = run(kind, a_src, layout, N, iters, rotate, grid)
                            except Exception as e:
                                print('%-5s %-7s %-6s %4d %6d %5d  ERROR %s' % (KIND[kind], SRC[a_src], LAY[layout], N, rotate, grid, e))
                                return
                            kk = 8 if kind == 0 else 16
                            print('%-5s %-7s %-6s %4d %6d %5d %10.1f %10.1f %12.0f' % (KIND[kind], SRC[a_src], LAY[layout], N, rotate, grid, m, mx,
                                                                                     128 * N * kk / m), flush=True)


if __name__ == '__main__':
    main()
