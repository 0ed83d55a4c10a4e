#!/usr/bin/env python
"""Where the end-to-end KRN step (bench.py `e2e`) loses time against the device-resident loop: times K steps of
(a) device batch, (b) + DevicePrefetcher H2D, (c) + loss D2H, (d) prefetcher with a device-resident 'loader' (no PCIe)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet     # noqa: E402
from speedplusbaseline_b200.optim import FusedAdamW                        # noqa: E402
from speedplusbaseline_b200.core.trainer import KRNTrainStep, DevicePrefetcher   # noqa: E402

dev = torch.device('cuda:0')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
model = KeypointRegressionNet(11, device=dev, seed=1)
model.train()
opt = FusedAdamW(model._store, model.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1, max_norm=1.0)
st = KRNTrainStep(model, opt)
h_img = torch.rand(48, 3, 224, 224).pin_memory()
h_tgt = torch.rand(48, 2, 11).pin_memory()
d_img, d_tgt = h_img.to(dev), h_tgt.to(dev)
for _ in range(5):
    st.step(d_img, d_tgt)
torch.cuda.synchronize()


def timed(name, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print('%-46s %.3f ms/step' % (name, e0.elapsed_time(e1) / K), flush=True)


def a():
    for _ in range(K):
        st.step(d_img, d_tgt)


def b():
    for db in DevicePrefetcher([(h_img, h_tgt)] * K, dev):
        st.step(*db)


host_loss = torch.empty(3).pin_memory()


def c():
    for db in DevicePrefetcher([(h_img, h_tgt)] * K, dev):
        host_loss.copy_(st.step(*db), non_blocking=True)


def d():
    for db in DevicePrefetcher([(d_img, d_tgt)] * K, dev):
        st.step(*db)


def e():      # graph replay only (static inputs already in place)
    for _ in range(K):
        st._graphs[0].replay()
        if st._graphs[1] is not None:
            st._graphs[1].replay()


# pure H2D bandwidth of the pinned batch
torch.cuda.synchronize()
buf = torch.empty_like(d_img)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    buf.copy_(h_img, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print('H2D of the pinned 28.9 MB batch: %.3f ms (%.1f GB/s)' % (e0.elapsed_time(e1) / 20, 28.9e-3 / (e0.elapsed_time(e1) / 20 / 1e3)), flush=True)

sampler = None
if len(sys.argv) > 2 and sys.argv[2] == 'sampler':
    sys.path.insert(0, ROOT)
    import bench
    sampler = bench.ClockSampler(0)
    sampler.start()
    print('clock sampler running', flush=True)

for name, fn in (('device batch', a), ('prefetcher, host batch (H2D)', b), ('prefetcher, host batch + loss D2H', c),
                 ('prefetcher, device batch (no PCIe)', d), ('graph replays only', e), ('device batch again', a)):
    timed(name, fn)
for name, fn in (('prefetcher, host batch + loss D2H (2)', c), ('device batch (2)', a), ('prefetcher, host batch + loss D2H (3)', c)):
    timed(name, fn)
if sampler:
    print(sampler.stop())
