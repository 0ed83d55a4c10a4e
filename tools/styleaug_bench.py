#!/usr/bin/env python
"""Time the style-augmentation forward (BASELINE.json configs[2] building block) at bs=48, 224x224:
whole call (CUDA events, median) and every launch of one call.  Synthetic Ghiasi weights (the real
checkpoint does not travel to the GPU box unless staged under baseline/_ref)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ghiasi as ogh, synth                       # noqa: E402  (weights only)
from speedplusbaseline_b200 import _lib as L, profiler       # noqa: E402
from speedplusbaseline_b200.styleaug.styleAugmentor import StyleAugmentor   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=48)
    ap.add_argument('--hw', type=int, default=224)
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    sd = synth.synth_state_dict(ogh.ghiasi_shapes(), 7)
    g = torch.Generator().manual_seed(1)
    cov = torch.randn(100, 100, generator=g)
    state = dict(ghiasi=sd, mean=torch.randn(1, 100, generator=g), cov=(cov @ cov.t() / 100).numpy(), base=torch.randn(100, generator=g))
    aug = StyleAugmentor(0.5, torch.device('cuda:0'), state=state)
    x = torch.rand(args.batch, 3, args.hw, args.hw, device='cuda')
    for _ in range(3):
        aug(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        aug(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gf = 15.434 * args.batch * (args.hw / 224.0) ** 2
    print('style-aug forward bs=%d %dx%d: %.3f ms  (%.0f img/s, %.1f TFLOP/s of the reference\'s %.0f GFLOP)'
          % (args.batch, args.hw, args.hw, ms, args.batch / ms * 1e3, gf / ms, gf))
    aug.use_graph = False                     # per-launch timing needs the un-captured call
    aug(x)
    with profiler.LaunchTimer() as lt:
        aug(x)
    rows = lt.rows()
    tot = sum(r[2] for r in rows)
    print('%d launches, %.1f us summed' % (len(rows), tot))
    for n, b, t, tag in rows:
        print('%-24s %9.1f us  %5.1f%%  %s' % (n, t, 100 * t / tot, tag[:60]))


if __name__ == '__main__':
    main()
