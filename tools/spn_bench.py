#!/usr/bin/env python
"""Time one SPN training step (BASELINE.json configs[4]: bs=32, 227x227) and every launch in it."""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speedplusbaseline_b200 import profiler                        # noqa: E402
from speedplusbaseline_b200.nets.spn import SpacecraftPoseNet       # noqa: E402
from speedplusbaseline_b200.spn_engine import SPNEngine             # noqa: E402
from speedplusbaseline_b200.optim import FusedAdamW                 # noqa: E402
from speedplusbaseline_b200.core.trainer import SPNTrainStep        # noqa: E402


def main():
    dev = torch.device('cuda:0')
    spn = SpacecraftPoseNet.__new__(SpacecraftPoseNet)
    torch.nn.Module.__init__(spn)
    spn.engine = SPNEngine(5000, device=dev)
    spn._register_store(spn.engine.store, spn.engine.key_order)
    spn.engine.store.params.normal_(0.0, 0.01)
    spn.train()
    opt = FusedAdamW(spn._store, spn.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=2)
    ss = SPNTrainStep(spn, opt, use_graph=False)
    B = 32
    x = torch.rand(B, 3, 227, 227, device=dev)
    yc = torch.zeros(B, 5000, device=dev)
    yc[:, :5] = 0.2
    for _ in range(3):
        ss.step(x, yc, yc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ss.step(x, yc, yc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('SPN train step bs=%d: %.3f ms (%.0f img/s)' % (B, ms, B / ms * 1e3))
    with profiler.LaunchTimer() as lt:
        ss.step(x, yc, yc)
    rows = lt.rows()
    tot = sum(r[2] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, b, t, tag in rows:
        agg[n][0] += 1
        agg[n][1] += t
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-24s %4d %9.1f us %5.1f%%' % (n, c, t, 100 * t / tot))
    print('--- top launches')
    for n, b, t, tag in sorted(rows, key=lambda r: -r[2])[:16]:
        print('%-24s %9.1f us  %s' % (n, t, tag[:70]))


if __name__ == '__main__':
    main()
