"""Oracle (test infrastructure): one training iteration of each reference loop.

krn_train_step   <- src/core/trainer.py:64-98  (fwd, zero_grad, backward, clip_grad_norm_ 1.0, AdamW)
dann_train_step  <- src/core/dann.py:74-100    (two forwards, BCE, backward, clip_grad_norm_ 1.0, AdamW)
spn_train_step   <- src/core/trainer.py:137-186 (CE + 10 CE, clip_grad_value_ 1.0, AdamW)

All state lives in a flat ``{key: tensor}`` dict; parameters are updated in
place; BN buffers are updated by the forward exactly as nn.BatchNorm2d does.
"""
import torch

from . import krn, revgrad, spn
from .optim import AdamWState, adamw_step, clip_grad_norm, clip_grad_value

HP = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01)   # README.md:78-85 recipe


def param_keys(sd):
    return [k for k in sd if krn.is_param(k)]


def _with_grad(sd, keys):
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None


def _finish(sd, keys, st, hp, clip):
    grads = []
    for k in keys:
        p = sd[k]
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
        p.requires_grad_(False)
    raw = {k: g.clone() for k, g in zip(keys, grads)}
    if clip == 'norm':
        total = clip_grad_norm(grads, 1.0)
    else:
        clip_grad_value(grads, 1.0)
        total = None
    adamw_step([sd[k] for k in keys], grads, st, **hp)
    for k in keys:
        sd[k].grad = None
    return raw, total


def krn_train_step(sd, st, images, target, hp=HP):
    keys = param_keys(sd)
    _with_grad(sd, keys)
    taps = {}
    loss, sm = krn.krn_forward(sd, images, target, train=True, taps=taps)
    loss.backward()
    raw, total = _finish(sd, keys, st, hp, 'norm')
    return dict(loss=float(loss), loss_x=sm['loss_x'], loss_y=sm['loss_y'], grads=raw,
                grad_norm=float(total), logits=taps['logits'].detach())


def dann_train_step(sd, st, source, label, target, alpha, hp=HP):
    keys = param_keys(sd)
    _with_grad(sd, keys)
    loss, lp, ls, lt = revgrad.dann_losses(sd, source, label, target, alpha)
    loss.backward()
    raw, total = _finish(sd, keys, st, hp, 'norm')
    return dict(loss=float(loss), loss_pose=float(lp), loss_source=float(ls),
                loss_target=float(lt), grads=raw, grad_norm=float(total))


def spn_train_step(sd, st, images, y_classes, y_weights, hp=HP, drop_p=0.0):
    keys = list(sd.keys())
    _with_grad(sd, keys)
    loss, lc, lr = spn.spn_loss(sd, images, y_classes, y_weights, train=True, drop_p=drop_p)
    loss.backward()
    raw, _ = _finish(sd, keys, st, hp, 'value')
    return dict(loss=float(loss), loss_class=float(lc), loss_regress=float(lr), grads=raw)


def new_state(sd):
    return AdamWState([sd[k] for k in param_keys(sd)])
