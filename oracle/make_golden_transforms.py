"""Generate tests/golden/transforms.npz by running the UNMODIFIED reference transforms
(/root/reference/src/datasets/transforms.py build_transforms :223-246) on seeded synthetic frames.
Build container only:  python -m oracle.make_golden_transforms

Per sample: a synthetic 8-bit frame (smooth + noise, grey replicated to RGB like `Image.open(..).convert('RGB')` of
a SPEED+ image), a bounding box and pixel keypoints go through the reference Compose with `torch.manual_seed(seed)`;
stored are the outputs and the seed, so tests can replay the same RNG stream through the oracle restatement.
"""
import os

import numpy as np

from oracle.make_golden import OUT, _shims

FRAME_HW = (300, 480)          # small frames keep the fixture small; the resampler is pinned against Pillow separately
CASES = [  # (seed, bbox xmin,xmax,ymin,ymax, p_aug, is_train, model)
    (1, (120., 300., 60., 200.), 1.0, True, 'krn'),
    (2, (10., 470., 5., 290.), 1.0, True, 'krn'),
    (3, (200., 260., 100., 180.), 0.5, True, 'krn'),
    (4, (0., 100., 0., 90.), 0.5, True, 'krn'),
    (5, (150., 400., 80., 250.), 0.0, False, 'krn'),
    (6, (33.5, 410.25, 20.75, 280.), 0.0, False, 'spn'),
]


def synth_frame(seed, hw=FRAME_HW):
    rng = np.random.default_rng(1000 + seed)
    H, W = hw
    yy, xx = np.mgrid[0:H, 0:W]
    g = 110 + 90 * np.sin(xx / 17.0 + seed) * np.cos(yy / 23.0) + rng.normal(0, 12, (H, W))
    return np.clip(g, 0, 255).astype(np.uint8)


def synth_keypoints(seed, bbox, K=11):
    rng = np.random.default_rng(2000 + seed)
    xmin, xmax, ymin, ymax = bbox
    return np.stack([rng.uniform(xmin, xmax, K), rng.uniform(ymin, ymax, K)]).astype(np.float32)


def main():
    _shims()
    import torch
    from PIL import Image
    from src.datasets.transforms import build_transforms
    out = {}
    for seed, bbox, p, is_train, model in CASES:
        grey = synth_frame(seed)
        data = Image.fromarray(grey).convert('RGB')
        size = (224, 224) if model == 'krn' else (227, 227)
        tf = build_transforms(model, size, p_aug=p, is_train=is_train)
        kp = synth_keypoints(seed, bbox)
        torch.manual_seed(seed)
        img, bb, k = tf(data, np.array(bbox, dtype=np.float32), kp.copy())
        out['img%d' % seed] = img.numpy()
        out['bbox%d' % seed] = np.asarray(bb, dtype=np.float32)
        out['kpt%d' % seed] = np.asarray(k, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, 'transforms.npz'), **out)
    print('written', os.path.join(OUT, 'transforms.npz'), os.path.getsize(os.path.join(OUT, 'transforms.npz')))


if __name__ == '__main__':
    main()
