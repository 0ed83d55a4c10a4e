"""Oracle (test infrastructure): the reference's per-sample input pipeline restated on the CPU
(SURVEY.md 8 row f1).  Never imported by the product path.

Follows /root/reference/src/datasets/transforms.py:
  RandomCrop :114-164, ResizeCrop :166-191, ToTensor :192-196, Rotate :38-57, Flip :59-72,
  BrightnessContrast :74-99, GaussianNoise :101-112, RandomApply :198-213, build_transforms :223-246
and Park2019KRNDataset.py:86-87 (`Image.open(..).convert('RGB')`).

Third-party arithmetic that is NOT under /root/reference: `T.resized_crop` on a PIL image = PIL crop + PIL
`Image.resize(size, BILINEAR)` (Pillow; requirements.txt:9 pins 8.4.0, this image has 12.2.0 -- the 8-bit
two-pass resampler of libImaging/Resample.c is the same algorithm in both).  Restated here from its published
source: per output index a window [xmin, xmax) around center = in0 + (i + .5) * scale with the triangle filter
stretched by max(scale, 1) (antialiasing), coefficients normalised in double and rounded to 22-bit fixed point,
each pass accumulated in int32 from 1 << 21, shifted by 22 and clipped to uint8; horizontal pass first (only over the
rows the vertical pass needs), a pass is skipped when it would not change the size.  `pil_resize_bilinear_u8` is pinned
against Pillow itself in tests/test_next_rows_cpu.py; the torch-side steps against the reference classes through
tests/golden/transforms.npz (oracle/make_golden_transforms.py).
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for the BILINEAR filter over the full extent.
    Returns (xmin [out], count [out], kk [out, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            if t < 0.0:
                t = -t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(n):
            if ww != 0.0:
                w[x] /= ww
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(src, xmin, cnt, kk, axis):
    """one resampling pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 array [H, W, C]."""
    src = np.moveaxis(src, axis, 0).astype(np.int64)
    out = np.empty((len(xmin),) + src.shape[1:], np.uint8)
    for i in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt[i]):
            acc += src[xmin[i] + x] * int(kk[i, x])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img, out_h, out_w):
    """`PIL.Image.resize((out_w, out_h), BILINEAR)` of a uint8 image [H, W, C]."""
    H, W = img.shape[:2]
    cur = img
    need_h, need_v = out_w != W, out_h != H
    if need_v:
        ymin, ycnt, ykk = pil_bilinear_coeffs(H, out_h)
    if need_h:
        xmin, xcnt, xkk = pil_bilinear_coeffs(W, out_w)
        first, last = (int(ymin[0]), int(ymin[-1] + ycnt[-1])) if need_v else (0, H)
        cur = _pass(cur[first:last], xmin, xcnt, xkk, 1)
        if need_v:
            ymin = ymin - first
    if need_v:
        cur = _pass(cur, ymin, ycnt, ykk, 0)
    return cur


# ---- transforms.py classes, with the random draws made explicit -------------------------------------
def random_crop_box(bbox, org_w, org_h, is_train, u=None):
    """RandomCrop :114-152.  bbox float32 [xmin, xmax, ymin, ymax]; u = the three torch.rand(1) draws (train).
    Returns integer (xmin, xmax, ymin, ymax) of the square-ish RoI clipped to the frame."""
    bbox = np.asarray(bbox, np.float32)
    xmin, xmax, ymin, ymax = bbox
    w, h = xmax - xmin, ymax - ymin
    x, y = xmin + w / 2.0, ymin + h / 2.0
    roi = max((w, h))
    if is_train:
        u0, u1, u2 = (torch.as_tensor(v, dtype=torch.float32).reshape(1) for v in u)
        roi = (1 + 0.5 * u0) * roi
        fx = 0.2 * (u1 * 2 - 1) * roi
        fy = 0.2 * (u2 * 2 - 1) * roi
    else:
        roi = (1 + 0.2) * roi
        fx = fy = 0
    x0 = max(0, int(x - roi / 2.0 + fx))
    x1 = min(org_w, int(x + roi / 2.0 + fx))
    y0 = max(0, int(y - roi / 2.0 + fy))
    y1 = min(org_h, int(y + roi / 2.0 + fy))
    return x0, x1, y0, y1


def resize_crop_box(bbox, org_w, org_h):
    """ResizeCrop :176-184 (SPN): the bounding box itself, clipped to the frame."""
    xmin, xmax, ymin, ymax = bbox
    return max(0, int(xmin)), min(org_w, int(xmax)), max(0, int(ymin)), min(org_h, int(ymax))


def crop_resize_to_tensor(frame_u8, box, out_hw):
    """T.resized_crop (PIL) + ToTensor: frame_u8 [H, W, C] uint8 -> float32 [C, out_h, out_w] in [0, 1]."""
    x0, x1, y0, y1 = box
    img = pil_resize_bilinear_u8(np.ascontiguousarray(frame_u8[y0:y1, x0:x1]), out_hw[0], out_hw[1])
    return torch.from_numpy(img).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


def crop_keypoints(keypts_pix, box):
    """RandomCrop :156-159: pixel keypoints [2, K] -> the crop's [0, 1] frame."""
    x0, x1, y0, y1 = box
    k = torch.as_tensor(np.asarray(keypts_pix), dtype=torch.float32).clone()
    k[0] = (k[0] - x0) / (x1 - x0)
    k[1] = (k[1] - y0) / (y1 - y0)
    return k


def apply_augment(image, keypts, rot=0, flip=0, bc=None, noise=None):
    """Rotate (rot = 1..3 quarter turns counter-clockwise), Flip (1 horizontal, 2 vertical), BrightnessContrast
    (bc = (a, b)), GaussianNoise (noise = std-scaled tensor) in the reference's order on a float image [C, H, W]
    and keypoints [2, K] in [0, 1]."""
    image, keypts = image.clone(), keypts.clone()
    if rot:
        image = torch.rot90(image, rot, (1, 2))          # == T.rotate(image, 90 * rot) on square inputs
        x, y = keypts[0].clone(), keypts[1].clone()
        if rot == 1:
            keypts[0], keypts[1] = y, 1.0 - x
        elif rot == 2:
            keypts[0], keypts[1] = 1.0 - x, 1.0 - y
        else:
            keypts[0], keypts[1] = 1.0 - y, x
    if flip == 1:
        image = image.flip(2)
        keypts[0] = 1.0 - keypts[0]
    elif flip == 2:
        image = image.flip(1)
        keypts[1] = 1.0 - keypts[1]
    if bc is not None:
        a, b = (torch.as_tensor(v, dtype=torch.float32).reshape(1) for v in bc)
        image = torch.clamp(a * image + b, 0, 1)
    if noise is not None:
        image = torch.clamp(image + noise, 0, 1)
    return image, keypts


def draw_augment(p, image_shape, gen=None):
    """RandomApply :207-211 over [Rotate, Flip, BrightnessContrast(.5..2, -25..25), GaussianNoise(25)]: consumes the torch
    RNG in the reference's order and returns the decisions dict(rot, flip, bc, noise)."""
    r = lambda *s: torch.rand(*s, generator=gen)
    d = dict(rot=0, flip=0, bc=None, noise=None)
    if r(1) < p:
        d['rot'] = int(torch.randint(1, 4, (1,), generator=gen))
    if r(1) < p:
        d['flip'] = 1 if r(1) < 0.5 else 2
    if r(1) < p:
        la = torch.tensor((0.5, 2.0)).log()
        lb = torch.tensor((-25, 25)) / 255
        loga = r(1) * (la[1] - la[0]) + la[0]
        a = loga.exp()
        b = r(1) * (lb[1] - lb[0]) + lb[0]
        d['bc'] = (a, b)
    if r(1) < p:
        d['noise'] = torch.randn(image_shape, dtype=torch.float32, generator=gen) * (25 / 255)
    return d
