#!/usr/bin/env python
"""B200SP_DEBUG_POISON=1: every activation / gradient buffer of the KRN engine starts as NaN; two eager steps at B=4, then report
which parameters / buffers hold NaN (a read-before-write shows up here)."""
import os, sys, torch
os.environ['B200SP_DEBUG_POISON'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
from speedplusbaseline_b200.optim import FusedAdamW
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = KeypointRegressionNet(11, device=dev, seed=100)
m.train()
opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
eng, st = m.engine, m._store
x, y = synth.synth_images(B, seed=10).to(dev), synth.synth_keypoints(B, seed=10).to(dev)
for step in range(2):
    st.grads.zero_()
    cx = eng.forward(x, y, train=True)
    torch.cuda.synchronize()
    for dn in ('Y', 'O'):
        for k, t in getattr(cx, dn).items():
            n = int(torch.isnan(t).sum())
            if n:
                print('step', step, 'after forward: NaN in', dn, k, tuple(t.shape), n, 'of', t.numel())
    eng.backward(cx)
    torch.cuda.synchronize()
    for dn in ('G', 'dO'):
        for k, t in getattr(cx, dn).items():
            n = int(torch.isnan(t).sum())
            if n:
                print('step', step, 'after backward: NaN in', dn, k, tuple(t.shape), n, 'of', t.numel())
    print('step', step, 'NaN in grads:', int(torch.isnan(st.grads).sum()), 'loss', float(eng.loss2[0]) if hasattr(eng, 'loss2') else '')
    opt.step()
torch.cuda.synchronize()
print('NaN in params:', int(torch.isnan(st.params).sum()), 'of', st.params.numel())
