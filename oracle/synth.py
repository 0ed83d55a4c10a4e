"""Oracle (test infrastructure): seeded synthetic weights and inputs.

Pretrained weights (ImageNet MobileNetV2, bvlc_alexnet.npy) cannot be fetched
offline, and a 22 MB random state dict cannot be committed, so every tensor is
*defined* by (seed, state_dict key, shape): the same function rebuilds
bit-identical weights in the build container (where they are loaded into the
unmodified reference modules to make tests/golden/) and on the GPU box (where
they are loaded into the oracle and into the CUDA path).  CPU torch RNG streams
are deterministic for a fixed torch version; tests/golden stores checksums of
the generated tensors so drift would be caught.
"""
import math
import zlib

import torch


def _gen(seed, key):
    g = torch.Generator(device='cpu')
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(seed, key, shape):
    g = _gen(seed, key)
    shape = tuple(shape)
    if key.endswith('num_batches_tracked'):
        return torch.zeros((), dtype=torch.int64)
    if key.endswith('running_mean'):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith('running_var'):
        return 0.5 + torch.rand(shape, generator=g)
    if len(shape) == 4 or len(shape) == 2:          # conv / linear weight: He-normal on fan_in
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return math.sqrt(2.0 / fan_in) * torch.randn(shape, generator=g)
    # 1-D: BN affine or a conv/linear bias.  BN weights sit next to a running_mean key;
    # callers pass bn=True through the key convention below.
    return None


def synth_state_dict(shapes, seed=2021):
    """shapes: OrderedDict key -> shape (oracle.*_shapes()).  Returns key -> tensor."""
    bn_prefixes = {k[:-len('.running_mean')] for k in shapes if k.endswith('.running_mean')}
    sd = {}
    for k, shp in shapes.items():
        t = synth_tensor(seed, k, shp)
        if t is None:
            g = _gen(seed, k)
            pre, leaf = k.rsplit('.', 1)
            if pre in bn_prefixes and leaf == 'weight':
                t = 0.5 + torch.rand(tuple(shp), generator=g)
            elif pre in bn_prefixes:
                t = 0.1 * torch.randn(tuple(shp), generator=g)
            else:
                t = 0.05 * torch.randn(tuple(shp), generator=g)
        sd[k] = t.contiguous()
    return sd


def synth_images(B, H=224, W=224, seed=2021, tag='images'):
    """[B,3,H,W] in [0,1): what ToTensor yields (datasets/transforms.py:192-196, no normalisation)."""
    return torch.rand((B, 3, H, W), generator=_gen(seed, '%s/%d/%d/%d' % (tag, B, H, W)))


def synth_keypoints(B, nk=11, seed=2021):
    """[B,2,nk] normalised keypoint coordinates (datasets/transforms.py:156-159)."""
    return torch.rand((B, 2, nk), generator=_gen(seed, 'kpts/%d/%d' % (B, nk)))


def synth_soft_targets(B, num_classes=5000, n_hot=5, seed=2021, tag='cls'):
    """Soft n-hot rows like SPNDataset.py:83-94 (n_hot classes, weights summing to 1)."""
    g = _gen(seed, 'soft/%s/%d/%d' % (tag, B, num_classes))
    t = torch.zeros(B, num_classes)
    for b in range(B):
        idx = torch.randperm(num_classes, generator=g)[:n_hot]
        w = torch.rand(n_hot, generator=g) + 0.1
        t[b, idx] = w / w.sum()
    return t


def checksum(t):
    """Order-sensitive float64 checksum used to pin regenerated tensors."""
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder_(97.0).add_(1.0)
    return float((t * w).sum())
