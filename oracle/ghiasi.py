"""Oracle (test infrastructure): style-augmentation net restatement.

Follows /root/reference/src/styleaug/ghiasi.py:6-135 (ConvInRelu :6-23,
UpsampleConvInRelu :26-62, ResidualBlock :65-103, Ghiasi :106-135) and the
embedding sampling / mixing of src/styleaug/styleAugmentor.py:39-68.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def ghiasi_shapes():
    d = OrderedDict()
    for i, (ci, co, k) in ((0, (3, 32, 9)), (1, (32, 64, 3)), (2, (64, 128, 3))):
        d['layers.%d.conv.weight' % i] = (co, ci, k, k)
        d['layers.%d.conv.bias' % i] = (co,)
    for i in range(3, 8):
        for j in (1, 2):
            d['layers.%d.conv%d.weight' % (i, j)] = (128, 128, 3, 3)
            d['layers.%d.conv%d.bias' % (i, j)] = (128,)
            for g in ('beta', 'gamma'):
                d['layers.%d.fc_%s%d.weight' % (i, g, j)] = (128, 100)
                d['layers.%d.fc_%s%d.bias' % (i, g, j)] = (128,)
    for i, (ci, co, k) in ((8, (128, 64, 3)), (9, (64, 32, 3)), (10, (32, 3, 9))):
        d['layers.%d.conv.weight' % i] = (co, ci, k, k)
        d['layers.%d.conv.bias' % i] = (co,)
        for g in ('beta', 'gamma'):
            d['layers.%d.fc_%s.weight' % (i, g)] = (co, 100)
            d['layers.%d.fc_%s.bias' % (i, g)] = (co,)
    return d


def _conv_in(x, w, b, stride):
    k = w.shape[2]
    x = F.pad(x, (k // 2,) * 4, mode='reflect')
    x = F.conv2d(x, w, b, stride)
    return F.instance_norm(x, eps=1e-5)          # InstanceNorm2d defaults: affine=False


def _cond(sd, p, sfx, style, x):
    beta = F.linear(style, sd['%s.fc_beta%s.weight' % (p, sfx)], sd['%s.fc_beta%s.bias' % (p, sfx)])
    gamma = F.linear(style, sd['%s.fc_gamma%s.weight' % (p, sfx)], sd['%s.fc_gamma%s.bias' % (p, sfx)])
    return gamma[:, :, None, None] * x + beta[:, :, None, None]


def ghiasi_forward(sd, x, style, taps=None):
    """ghiasi.py:125-135.  x [B,3,H,W] in [0,1], style [B,100] -> [B,3,H,W] in (0,1)."""
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t
    for i, s in ((0, 1), (1, 2), (2, 2)):
        p = 'layers.%d' % i
        x = tap(p, F.relu(_conv_in(x, sd[p + '.conv.weight'], sd[p + '.conv.bias'], s)))
    for i in range(3, 8):
        p = 'layers.%d' % i
        y = _conv_in(x, sd[p + '.conv1.weight'], sd[p + '.conv1.bias'], 1)
        y = F.relu(_cond(sd, p, '1', style, y))
        y = _conv_in(y, sd[p + '.conv2.weight'], sd[p + '.conv2.bias'], 1)
        y = _cond(sd, p, '2', style, y)
        x = tap(p, x + y)
    for i, up, act in ((8, 2, True), (9, 2, True), (10, None, False)):
        p = 'layers.%d' % i
        if up:
            x = F.interpolate(x, scale_factor=up)          # nn.Upsample default: nearest
        x = _conv_in(x, sd[p + '.conv.weight'], sd[p + '.conv.bias'], 1)
        x = _cond(sd, p, '', style, x)
        if act:
            x = F.relu(x)
        tap(p, x)
    return torch.sigmoid(x)


def style_matrix(cov):
    """styleAugmentor.py:39-42: A = U * sqrt(S) from the SVD of the covariance."""
    u, s, _ = np.linalg.svd(np.asarray(cov, dtype=np.float64))
    return torch.tensor(np.matmul(u, np.diag(s ** 0.5))).float()


def mix_embedding(noise, A, mean, base, alpha):
    """styleAugmentor.py:44-64 with the randn draw supplied by the caller:
    z = noise @ A^T + mean;  e = alpha*z + (1-alpha)*base."""
    z = torch.mm(noise, A.transpose(1, 0)) + mean
    return alpha * z + (1 - alpha) * base
