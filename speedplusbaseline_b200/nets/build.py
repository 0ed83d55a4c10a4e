"""Model / optimizer factory with the reference's contract (/root/reference/src/nets/build.py:39-78).

get_model dispatches on cfg.model_name / cfg.dann exactly like the reference; get_optimizer returns the
fused flat-buffer AdamW for `--optimizer adamw` (the north-star path; clip mode follows the loop that
will drive it: global-norm 1.0 for KRN/DANN, trainer.py:97 / dann.py:99, value 1.0 for SPN,
trainer.py:184).  sgd / rmsprop / adam map to the same flat-buffer design
(optim.FusedSGD / FusedRMSprop / FusedAdam -> b200sp_optim_step) with the reference's argument mapping
(`cfg.momentum` is SGD's momentum, RMSprop's alpha and Adam's beta1, build.py:63-71).
"""
import logging

import torch

from .. import _lib as L
from ..optim import FusedAdam, FusedAdamW, FusedRMSprop, FusedSGD
from .park2019 import KeypointRegressionNet
from .revgrad import RevGrad

logger = logging.getLogger(__name__)


def _dtype(cfg):
    """storage dtype: fp32.  (B200SP_ENABLE_BF16=1 selects the experimental bf16-STORAGE engine instead)"""
    return L.BF16 if getattr(cfg, 'fp16', False) and L.BF16_ENABLED else L.F32


def _tf32(cfg):
    """`--use_fp16` (config.py:39; reference: torch.cuda.amp autocast + GradScaler, trainer.py:73-94) selects the mixed-precision
    mode of this path: the 1x1-convolution GEMMs run SINGLE-pass TF32 on the tensor cores -- operands rounded to fp16's 10-bit
    mantissa, fp32 exponent range, fp32 accumulation -- while storage, BatchNorm statistics, depthwise stencils, loss and the
    optimizer stay fp32.  The exponent range is fp32's, so the reference's loss scaling has nothing to protect: the GradScaler
    the CLI creates is accepted by the epoch loops and left at scale 1 (tests/test_krn_tf32_gpu.py)."""
    return bool(getattr(cfg, 'fp16', False)) and not L.BF16_ENABLED


def get_model(cfg):
    assert cfg.model_name == 'krn' or cfg.model_name == 'spn', \
        'Model name must be either krn or spn'
    device = getattr(cfg, 'device', None)
    if not cfg.dann:
        if cfg.model_name == 'krn':
            model = KeypointRegressionNet(cfg.num_keypoints, device=device, dtype=_dtype(cfg), tf32_gemm=_tf32(cfg))
            logger.info('KRN created' + (' (--use_fp16: single-pass TF32 GEMMs, fp32 storage)' if _tf32(cfg) else ''))
        else:
            from .spn import SpacecraftPoseNet
            model = SpacecraftPoseNet(cfg.num_classes, pretrain=True, device=device)
            logger.info('SPN created')
    else:
        model = RevGrad(cfg.num_keypoints, device=device)
        logger.info('RevGrad created with {}'.format(cfg.model_name))
    n = sum(p.numel() for p in model.parameters())
    logger.info('   - Number of total parameters:     {:,}'.format(n))
    logger.info('   - Number of trainable parameters: {:,}'.format(n))
    return model


def get_optimizer(cfg, model):
    param = filter(lambda p: p.requires_grad, model.parameters())
    clip_mode = 2 if cfg.model_name == 'spn' and not cfg.dann else 1
    clip = dict(clip_mode=clip_mode, max_norm=1.0, clip_value=1.0)
    if cfg.optimizer == 'sgd':
        optimizer = FusedSGD(model._store, param, lr=cfg.lr, momentum=cfg.momentum, weight_decay=cfg.weight_decay, **clip)
    elif cfg.optimizer == 'rmsprop':
        optimizer = FusedRMSprop(model._store, param, lr=cfg.lr, alpha=cfg.momentum, weight_decay=cfg.weight_decay, **clip)
    elif cfg.optimizer == 'adam':
        optimizer = FusedAdam(model._store, param, lr=cfg.lr, betas=(cfg.momentum, 0.999), weight_decay=cfg.weight_decay, **clip)
    elif cfg.optimizer == 'adamw':
        optimizer = FusedAdamW(model._store, param, lr=cfg.lr, betas=(cfg.momentum, 0.999),
                               weight_decay=cfg.weight_decay, clip_mode=clip_mode, max_norm=1.0, clip_value=1.0)
    else:
        raise AssertionError('unknown optimizer %r' % cfg.optimizer)
    logger.info('Optimizer created: {}'.format(cfg.optimizer))
    return optimizer
