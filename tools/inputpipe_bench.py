"""Time the device input pipeline on SPEED+-sized frames: whole call (host decisions + launches) and the
three kernels alone (fixed decisions, CUDA events).  python tools/inputpipe_bench.py [--reps N] [--batch B]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from speedplusbaseline_b200.datasets.transforms import build_transforms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--batch', type=int, default=48)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    B = a.batch
    tf = build_transforms('krn', (224, 224), p_aug=0.5, is_train=True, device=dev, generator=torch.Generator().manual_seed(1))
    frames = torch.randint(0, 256, (B, 1200, 1920), dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(1)
    cx, cy, sz = rng.uniform(500, 1400, B), rng.uniform(400, 800, B), rng.uniform(150, 700, B)
    bbox = np.stack([cx - sz / 2, cx + sz / 2, cy - sz / 2, cy + sz / 2], 1).astype(np.float32)
    kp = np.zeros((B, 2, 11), np.float32)
    augs, boxes = tf.sample(bbox, 1920, 1200)
    crop_px = sum((b[1] - b[0]) * (b[3] - b[2]) for b in boxes)
    tmp_px = sum((b[3] - b[2]) * 224 for b in boxes)
    alg = crop_px + 2 * tmp_px + B * 3 * 224 * 224 * 4              # bytes: crop read, temp written + read, fp32 NCHW written
    for _ in range(3):
        tf.apply(frames, augs, kp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        tf.apply(frames, augs, kp)
    e1.record()
    torch.cuda.synchronize()
    ms_apply = e0.elapsed_time(e1) / a.reps
    t0 = time.perf_counter()
    for _ in range(a.reps):
        tf(frames, bbox, kp)
    torch.cuda.synchronize()
    ms_call = (time.perf_counter() - t0) / a.reps * 1e3
    t0 = time.perf_counter()
    for _ in range(a.reps):
        tf.sample(bbox, 1920, 1200)
    ms_host = (time.perf_counter() - t0) / a.reps * 1e3
    print('# input pipeline, %d grey 1200x1920 frames -> [%d,3,224,224] fp32 (status %d)' % (B, B, tf.status()))
    print('apply() fixed decisions, CUDA events : %.3f ms  (%.0f images/s, %.1f MB algorithmic -> %.0f GB/s)'
          % (ms_apply, B / ms_apply * 1e3, alg / 1e6, alg / ms_apply / 1e6))
    print('full call incl. host sampling, wall  : %.3f ms  (%.0f images/s)' % (ms_call, B / ms_call * 1e3))
    print('host decision sampling alone         : %.3f ms' % ms_host)


if __name__ == '__main__':
    main()
