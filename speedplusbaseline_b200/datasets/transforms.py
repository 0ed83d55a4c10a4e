"""Batched, on-device version of the reference's per-sample transform stack.

`build_transforms(model_name, input_size, p_aug, is_train)` has the reference's signature
(/root/reference/src/datasets/transforms.py:223-246) but returns a `DeviceTransforms` that is called ONCE PER BATCH:

    images, bbox, keypts = tf(frames_u8, bbox, keypts_pix)

    frames_u8  uint8 [B,H,W] or [B,H,W,C] (C in {1,3}), device or pinned host -- the decoded camera frames
               (Park2019KRNDataset.py:86 `Image.open(..).convert('RGB')`; grey frames are replicated to RGB on the device)
    bbox       float32 [B,4] (xmin, xmax, ymin, ymax) pixels, host
    keypts_pix float32 [B,2,K] pixels, host (or None)
    -> images  float32 [B,3,h,w] in [0,1] on the device (what KRNTrainStep.step / the engines consume),
       bbox    float32 [B,4] host (KRN: the sampled RoI, RandomCrop :154; SPN: unchanged, ResizeCrop :189),
       keypts  float32 [B,2,K] device, normalised to the RoI and rotated / flipped with the image

The random DECISIONS (RoI enlargement / shift, which augmentations fire, rotation, flip direction, contrast / brightness
factors) are sampled on the host from torch's RNG with the reference's distributions and order (RandomCrop :136-142,
RandomApply :207-211, Rotate :41, Flip :61, BrightnessContrast :89-94); the pixel work -- crop, Pillow-exact bilinear
resize, ToTensor, rotation, flip, a*x+b, Gaussian noise -- runs in three libb200sp launches (csrc/inputpipe.cu).  The noise
itself is drawn by a counter-based device generator (statistical parity with GaussianNoise :110, not stream parity).
No CPU fallback: without the CUDA library this module raises.
"""
import ctypes as C
import math

import numpy as np
import torch

from .. import _lib as L


def dest_index(sy, sx, oh, ow, rot, flip):
    """Host mirror of inputpipe.cu:dest_index -- where source pixel (sy, sx) lands after torch.rot90(img, rot, (1, 2)) and the
    flip (1 horizontal, 2 vertical).  Unit-tested against torch on the CPU."""
    if rot == 1:
        i, j = ow - 1 - sx, sy
    elif rot == 2:
        i, j = oh - 1 - sy, ow - 1 - sx
    elif rot == 3:
        i, j = sx, oh - 1 - sy
    else:
        i, j = sy, sx
    if flip == 1:
        j = ow - 1 - j
    elif flip == 2:
        i = oh - 1 - i
    return i, j


def sample_crop_box(bbox, org_w, org_h, is_train, gen=None):
    """RandomCrop :124-152 for one sample: float32 torch arithmetic in the reference's order, three torch.rand(1) draws."""
    bbox = np.asarray(bbox, np.float32)
    xmin, xmax, ymin, ymax = bbox
    w, h = xmax - xmin, ymax - ymin
    x, y = xmin + w / 2.0, ymin + h / 2.0
    roi = max((w, h))
    if is_train:
        roi = (1 + 0.5 * torch.rand(1, generator=gen)) * roi
        fx = 0.2 * (torch.rand(1, generator=gen) * 2 - 1) * roi
        fy = 0.2 * (torch.rand(1, generator=gen) * 2 - 1) * roi
    else:
        roi = (1 + 0.2) * roi
        fx = fy = 0
    return (max(0, int(x - roi / 2.0 + fx)), min(org_w, int(x + roi / 2.0 + fx)),
            max(0, int(y - roi / 2.0 + fy)), min(org_h, int(y + roi / 2.0 + fy)))


def sample_augment(p, gen=None):
    """RandomApply :207-211 over [Rotate, Flip, BrightnessContrast(alpha .5..2, beta -25..25), GaussianNoise(25)] for one
    sample.  Returns (rot, flip, bc, a, b, noise_std); the noise tensor itself is NOT drawn here (device generator)."""
    r = lambda: torch.rand(1, generator=gen)
    rot = flip = bc = 0
    a, b, std = 1.0, 0.0, 0.0
    if r() < p:
        rot = int(torch.randint(1, 4, (1,), generator=gen))
    if r() < p:
        flip = 1 if r() < 0.5 else 2
    if r() < p:
        la, lb = torch.tensor((0.5, 2.0)).log(), torch.tensor((-25, 25)) / 255
        a = float((r() * (la[1] - la[0]) + la[0]).exp())
        b = float(r() * (lb[1] - lb[0]) + lb[0])
        bc = 1
    if r() < p:
        std = 25 / 255
    return rot, flip, bc, a, b, std


def sample_batch(bbox, org_w, org_h, model_name, is_train, p, gen=None, u=None):
    """Decisions for a whole batch from ONE block of uniforms u [B, 11] (drawn here unless given):
    columns 0-2 RoI enlargement / x shift / y shift (RandomCrop :136-142); 3 rotate?, 4 which quarter turn (Rotate :41:
    randint(1,4) == 1 + floor(3u)); 5 flip?, 6 direction (Flip :61); 7 brightness-contrast?, 8-9 its alpha / beta draws
    (:89-94); 10 noise? (RandomApply :209).  Float32 torch arithmetic in the reference's operation order, so a row given the
    reference's own uniforms reproduces `sample_crop_box` and the brightness / contrast factors of `sample_augment` exactly
    (tests/test_next_rows_cpu.py)."""
    B = bbox.shape[0]
    if u is None:
        u = torch.rand(B, 11, generator=gen)
    u = u.to(torch.float32)
    bb = torch.from_numpy(np.ascontiguousarray(bbox, dtype=np.float32))
    xmin, xmax, ymin, ymax = bb[:, 0], bb[:, 1], bb[:, 2], bb[:, 3]
    if model_name == 'krn':
        w, h = xmax - xmin, ymax - ymin
        x, y = xmin + w / 2.0, ymin + h / 2.0
        roi = torch.maximum(w, h)
        if is_train:
            roi = (1 + 0.5 * u[:, 0]) * roi
            fx = 0.2 * (u[:, 1] * 2 - 1) * roi
            fy = 0.2 * (u[:, 2] * 2 - 1) * roi
        else:
            roi = (1 + 0.2) * roi
            fx = fy = torch.zeros(B)
        x0 = (x - roi / 2.0 + fx).to(torch.int64).clamp_(min=0)           # int(): truncation toward zero
        x1 = (x + roi / 2.0 + fx).to(torch.int64).clamp_(max=org_w)
        y0 = (y - roi / 2.0 + fy).to(torch.int64).clamp_(min=0)
        y1 = (y + roi / 2.0 + fy).to(torch.int64).clamp_(max=org_h)
    else:                                                                   # ResizeCrop :176-184
        x0, x1 = xmin.to(torch.int64).clamp_(min=0), xmax.to(torch.int64).clamp_(max=org_w)
        y0, y1 = ymin.to(torch.int64).clamp_(min=0), ymax.to(torch.int64).clamp_(max=org_h)
    box = torch.stack([x0, x1, y0, y1], 1).tolist()
    rot, flip, bc = [0] * B, [0] * B, [0] * B
    a, b, std, seed = [1.0] * B, [0.0] * B, [0.0] * B, list(range(1, B + 1))
    if is_train and model_name == 'krn':
        rot = torch.where(u[:, 3] < p, 1 + (u[:, 4] * 3).to(torch.int64).clamp_(max=2), torch.zeros(B, dtype=torch.int64)).tolist()
        flip = torch.where(u[:, 5] < p, torch.where(u[:, 6] < 0.5, 1, 2), torch.zeros(B, dtype=torch.int64)).tolist()
        la, lb = torch.tensor((0.5, 2.0)).log(), torch.tensor((-25, 25)) / 255
        on = u[:, 7] < p
        bc = on.to(torch.int64).tolist()
        a = torch.where(on, (u[:, 8] * (la[1] - la[0]) + la[0]).exp(), torch.ones(B)).tolist()
        b = torch.where(on, u[:, 9] * (lb[1] - lb[0]) + lb[0], torch.zeros(B)).tolist()
        std = torch.where(u[:, 10] < p, torch.full((B,), 25 / 255), torch.zeros(B)).tolist()
        seed = torch.randint(0, 2 ** 31 - 1, (B,), generator=gen).tolist()
    return dict(box=[tuple(r) for r in box], rot=rot, flip=flip, bc=bc, a=a, b=b, std=std, seed=seed)


class DeviceTransforms:
    def __init__(self, model_name, input_size, p_aug=0.5, is_train=True, device=None, generator=None):
        assert model_name in ('krn', 'spn')
        L.require_cuda()
        self.model_name, self.p, self.is_train = model_name, p_aug, is_train
        self.oh, self.ow = int(input_size[0]), int(input_size[1])
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.gen = generator
        self._scratch = {}
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._seed = 0

    # decisions for a batch -> list of L.Aug
    def sample(self, bbox, org_w, org_h):
        """one batch of decisions.  Vectorised (`sample_batch`): the per-sample functions above follow the reference's RNG
        draw ORDER (kept for the parity tests); here every uniform of the batch comes from one torch.rand call and goes through
        the same float32 formulas -- same distributions, ~50x less host time at bs=48 (the reference's DataLoader workers each
        have their own stream anyway, so there is no batch-level stream to match)."""
        dec = sample_batch(np.asarray(bbox, np.float32), org_w, org_h, self.model_name, self.is_train, self.p, self.gen)
        augs = []
        for i in range(len(dec['box'])):
            x0, x1, y0, y1 = dec['box'][i]
            augs.append(L.Aug(x0, x1, y0, y1, dec['rot'][i], dec['flip'][i], dec['bc'][i], dec['a'][i], dec['b'][i],
                              dec['std'][i], dec['seed'][i], 0))
        return augs, dec['box']

    def _buf(self, name, numel, dtype):
        t = self._scratch.get(name)
        if t is None or t.numel() < numel:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self._scratch[name] = t
        return t

    def apply(self, frames, augs, keypts_pix=None, normalize_kpts=True):
        """run the device pipeline with explicit per-image decisions (list of L.Aug)."""
        if frames.dtype != torch.uint8:
            raise TypeError('frames must be uint8 (decoded camera frames)')
        if frames.dim() == 3:
            frames = frames.unsqueeze(-1)
        B, H, W, Cc = frames.shape
        if len(augs) != B:
            raise ValueError('one b200sp_aug per frame')
        fr = frames.to(self.device, non_blocking=True).contiguous()
        for g in augs:
            if not (0 <= g.x0 < g.x1 <= W and 0 <= g.y0 < g.y1 <= H):
                raise ValueError('crop box (%d,%d,%d,%d) outside the %dx%d frame' % (g.x0, g.x1, g.y0, g.y1, W, H))
        max_w, max_h = max(g.x1 - g.x0 for g in augs), max(g.y1 - g.y0 for g in augs)
        ks = 2 * int(math.ceil(max(max_w / self.ow, max_h / self.oh, 1.0))) + 1
        arr = (L.Aug * B)(*augs)
        aug_d = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device, non_blocking=True)
        coef = self._buf('coef', B * 2 * max(self.oh, self.ow) * (2 + ks), torch.int32)
        tmp = self._buf('tmp', B * max_h * self.ow * Cc, torch.uint8)
        out = torch.empty(B, 3, self.oh, self.ow, device=self.device)
        odd = int(any(g.rot in (1, 3) for g in augs))
        sp = L.stream_ptr()
        L.call('b200sp_input_pipeline', fr.data_ptr(), B, H, W, Cc, aug_d.data_ptr(), coef.data_ptr(), tmp.data_ptr(), max_h, ks,
               out.data_ptr(), self.oh, self.ow, odd, self._status.data_ptr(), sp)
        kout = None
        if keypts_pix is not None:
            kin = torch.as_tensor(np.asarray(keypts_pix), dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
            kout = torch.empty_like(kin)
            L.call('b200sp_kpt_augment', kin.data_ptr(), aug_d.data_ptr(), kout.data_ptr(), B, kin.shape[2], 1 if normalize_kpts else 0, sp)
        self._keep = (fr, aug_d)            # the launches above are asynchronous: keep their inputs alive until the next call
        return out, kout

    def status(self):
        """device-side guard flags of the calls so far (0 = every box fitted the scratch sizing); synchronises."""
        return int(self._status.item())

    def __call__(self, frames, bbox, keypts_pix=None):
        H, W = frames.shape[1], frames.shape[2]
        augs, boxes = self.sample(bbox, W, H)
        images, kout = self.apply(frames, augs, keypts_pix, normalize_kpts=self.model_name == 'krn')
        if self.model_name == 'krn':
            bbox_out = torch.tensor(boxes, dtype=torch.float32)
        else:
            bbox_out = torch.as_tensor(np.asarray(bbox), dtype=torch.float32)
        return images, bbox_out, kout


def build_transforms(model_name, input_size, p_aug=0.5, is_train=True, device=None, generator=None):
    return DeviceTransforms(model_name, input_size, p_aug, is_train, device, generator)
