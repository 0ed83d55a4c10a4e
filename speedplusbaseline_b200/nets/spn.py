"""SpacecraftPoseNet -- same constructor / forward contract as /root/reference/src/nets/spn.py:50-143
(`SpacecraftPoseNet(num_classes, keep_prob=0.5, pretrain=True)(x) -> (c, r)` logits), executed by libb200sp
kernels (spn_engine.SPNEngine), plus the soft-target cross entropy of spn.py:37-48."""
import logging
import math
import os

import numpy as np
import torch

from .. import _lib as L
from ..spn_engine import SPNEngine
from .park2019 import EngineModule, default_init

logger = logging.getLogger(__name__)


def softmax_cross_entropy_with_logits(logits, target, reduction='mean'):
    """spn.py:37-48 (used by loops that compute the loss outside the fused step)."""
    loss = -torch.sum(target.detach() * torch.nn.functional.log_softmax(logits, dim=1), dim=1)
    if reduction == 'mean':
        return loss.mean()
    return loss.sum() if reduction == 'sum' else loss


class _SPNPass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, fwd, bwd):
        ctx.bwd = bwd
        c, r = fwd()
        return c, r

    @staticmethod
    def backward(ctx, gc, gr):
        ctx.bwd(gc, gr)
        return None, None, None


class SpacecraftPoseNet(EngineModule):
    def __init__(self, num_classes, keep_prob=0.5, pretrain=True, device=None, seed=None):
        super().__init__()
        self.num_classes = num_classes
        self.regress_size = num_classes
        self.keep_prob = keep_prob
        self.engine = SPNEngine(num_classes, device=device)
        self.engine.drop_p = 0.5                      # spn.py:81: nn.Dropout(0.5) regardless of keep_prob
        self._register_store(self.engine.store, self.engine.key_order)
        default_init(self.engine.store, seed, kaiming_prefixes=())
        if pretrain:
            self.load_weights('checkpoints/pretrained/bvlc_alexnet.npy')

    def load_weights(self, weight_path):
        """spn.py:101-123: first five conv layers from the Caffe AlexNet dump ([H,W,Cin,Cout] -> [Cout,Cin,H,W])."""
        if not os.path.exists(weight_path):
            logger.warning('   - %s not found: SPN convolutions keep their random initialisation', weight_path)
            return
        from ..importers import load_alexnet_npy
        load_alexnet_npy(self, weight_path)

    def forward(self, x):
        eng = self.engine
        x = x.contiguous().float()
        if not torch.is_grad_enabled() or not self.training:
            c, r = eng.forward(x, train=self.training)
            return c.clone(), r.clone()

        def fwd():
            c, r = eng.forward(x, train=True)
            return c.clone(), r.clone()

        def bwd(gc, gr):
            self.rebind_grads()
            B = x.shape[0]
            eng._buf('dz_c', (B, eng.nc)).copy_(gc if gc is not None else 0)
            eng._buf('dz_r', (B, eng.nc)).copy_(gr if gr is not None else 0)
            eng.backward()

        return _SPNPass.apply(self._plist[0], fwd, bwd)
