#!/usr/bin/env python
"""How deterministic is the two-shard single-process reference of tests/test_ddp_gpu.py (run twice), and does the graph step
equal the eager step on one GPU?  Prints relative differences (max |a-b| / max |b|) and where the largest one sits."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import synth                                                  # noqa: E402
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet   # noqa: E402
from speedplusbaseline_b200.optim import FusedAdamW                       # noqa: E402
from speedplusbaseline_b200.core.trainer import KRNTrainStep              # noqa: E402
import test_ddp_gpu as T                                                   # noqa: E402


def rel(a, b):
    d = (a.double() - b.double()).abs()
    i = int(d.argmax())
    return float(d.max() / b.double().abs().max()), i, float(a[i]), float(b[i])


r1 = T._single_process_reference('krn', 2)
r2 = T._single_process_reference('krn', 2)
print('reference run-to-run:', rel(r1, r2))


def dist_stats(a, b, tag):
    d = (a.double() - b.double()).abs()
    print(tag, 'max %.3e  frac>1e-4 %.3e  frac>1e-5 %.3e  q99.9 %.3e  median %.3e  n %d' % (
        float(d.max()), float((d > 1e-4).double().mean()), float((d > 1e-5).double().mean()),
        float(torch.quantile(d[::7].float(), 0.999)), float(d.median()), d.numel()))


dist_stats(r1, r2, 'reference run-to-run |diff|:')
dev = torch.device('cuda:0')
outs = {}
for mode in ('eager', 'graph'):
    m = KeypointRegressionNet(11, device=dev, seed=100)
    m.train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    x, y = synth.synth_images(4, seed=10).to(dev), synth.synth_keypoints(4, seed=10).to(dev)
    st = KRNTrainStep(m, opt, use_graph=mode == 'graph')
    for _ in range(2):
        st.step(x, y)
    torch.cuda.synchronize()
    outs[mode] = m._store.params.cpu().clone()
print('graph vs eager (1 GPU, 2 steps):', rel(outs['graph'], outs['eager']))
dist_stats(outs['graph'], outs['eager'], 'graph vs eager |diff|:')
m = KeypointRegressionNet(11, device=dev, seed=100)
names = []
off = 0
for k, p in m.named_parameters():
    names.append((off, off + p.numel(), k))
    off += p.numel()
for tag, (e, i, a, b) in (('ref', rel(r1, r2)), ('graph', rel(outs['graph'], outs['eager']))):
    for lo, hi, k in names:
        if lo <= i < hi:
            print(tag, 'largest difference in', k, 'index', i - lo, a, b)
