"""Oracle (test infrastructure): DANN / RevGrad restatement.

Follows /root/reference/src/nets/revgrad.py:36-96 (GradientReversalFunction
:36-56, RevGrad :58-96) and the DANN step of src/core/dann.py:74-100.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .krn import krn_shapes, krn_logits, krn_loss


class _GRL(torch.autograd.Function):
    # revgrad.py:46-56: forward copy, backward -lambda * g
    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = lambda_
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return -g.new_tensor(ctx.lambda_) * g, None


def revgrad_shapes(num_keypoints=11):
    d = krn_shapes(num_keypoints, 'net.')
    d['domain_classifier.0.weight'] = (1280, 320, 1, 1)   # revgrad.py:75-80
    d['domain_classifier.0.bias'] = (1280,)
    d['domain_classifier.3.weight'] = (1, 1280, 1, 1)
    d['domain_classifier.3.bias'] = (1,)
    return d


def domain_head(sd, feature, alpha):
    h = _GRL.apply(feature, alpha)
    h = F.relu(F.conv2d(h, sd['domain_classifier.0.weight'], sd['domain_classifier.0.bias']))
    h = F.avg_pool2d(h, 7)
    h = F.conv2d(h, sd['domain_classifier.3.weight'], sd['domain_classifier.3.bias'])
    return h.squeeze()


def revgrad_forward(sd, x, y=None, alpha=None, train=True):
    """revgrad.py:82-96.  Returns (out1, domain_logits) when alpha is given."""
    feature, logits = krn_logits(sd, x, train, 'net.')
    if y is not None:
        loss, lx, ly = krn_loss(logits, y)
        out1 = (loss, {'loss_x': float(lx.detach()), 'loss_y': float(ly.detach())})
    else:
        out1 = (logits[:, 0::2], logits[:, 1::2])
    if alpha is None:
        return out1
    return out1, domain_head(sd, feature, alpha)


def dann_alpha(idx, epoch, n_batches, max_epochs):
    """dann.py:77-78."""
    p = float(idx + epoch * n_batches) / max_epochs / n_batches
    return 2. / (1. + np.exp(-10 * p)) - 1


def dann_losses(sd, source, label, target, alpha):
    """dann.py:81-95: returns (total, pose, dom_src, dom_tgt)."""
    B = source.shape[0]
    (loss_pose, _), dsrc = revgrad_forward(sd, source, label, alpha, True)
    l_src = F.binary_cross_entropy_with_logits(dsrc, torch.ones(B), reduction='mean')
    _, dtgt = revgrad_forward(sd, target, None, alpha, True)
    l_tgt = F.binary_cross_entropy_with_logits(dtgt, torch.zeros(B), reduction='mean')
    return loss_pose + l_src + l_tgt, loss_pose, l_src, l_tgt
