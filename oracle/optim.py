"""Oracle (test infrastructure): clip + AdamW restated as plain tensor arithmetic.

The reference calls torch for these (src/nets/build.py:72-74 ->
torch.optim.AdamW(lr, betas=(momentum, 0.999), weight_decay), eps 1e-8,
amsgrad off; src/core/trainer.py:90,97 and src/core/dann.py:99 ->
clip_grad_norm_(params, 1.0); trainer.py:177,184 -> clip_grad_value_(params,
1.0)).  The published algorithms restated here: torch/optim/adam.py
(_single_tensor_adam, decoupled decay branch) and
torch/nn/utils/clip_grad.py (clip_coef = max_norm / (total_norm + 1e-6),
clamped to 1).
"""
import math

import torch


def clip_grad_norm(grads, max_norm=1.0):
    """grads: list of tensors, scaled in place; returns the total L2 norm."""
    total = torch.linalg.vector_norm(
        torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def clip_grad_value(grads, clip=1.0):
    for g in grads:
        g.clamp_(-clip, clip)


class AdamWState:
    def __init__(self, params):
        self.step = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]


def adamw_step(params, grads, st, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01):
    """One AdamW step in place on ``params`` (list of tensors)."""
    st.step += 1
    bc1 = 1 - beta1 ** st.step
    bc2 = 1 - beta2 ** st.step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, st.m, st.v):
        p.mul_(1 - lr * wd)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


# ---- the other `--optimizer` choices (src/nets/build.py:63-71) ------------------------------------
# torch.optim.SGD(lr, momentum=cfg.momentum, weight_decay), RMSprop(lr, alpha=cfg.momentum, weight_decay)
# and Adam(lr, betas=(cfg.momentum, 0.999), weight_decay); published algorithms: torch/optim/sgd.py
# (_single_tensor_sgd: dampening 0, nesterov off), torch/optim/rmsprop.py (_single_tensor_rmsprop:
# eps 1e-8, momentum 0, not centered), torch/optim/adam.py (_single_tensor_adam, coupled decay).
# All three add weight_decay * param to the gradient first (L2, not decoupled).
class FlatState:
    def __init__(self, params):
        self.step = 0
        self.s1 = [torch.zeros_like(p) for p in params]      # momentum_buffer | square_avg | exp_avg
        self.s2 = [torch.zeros_like(p) for p in params]      # exp_avg_sq (Adam)


def sgd_step(params, grads, st, lr=1e-3, momentum=0.9, wd=0.0):
    st.step += 1
    for p, g, buf in zip(params, grads, st.s1):
        g = g.add(p, alpha=wd)
        if st.step == 1:
            buf.copy_(g)                                       # torch: buf = clone(grad) on the first step
        else:
            buf.mul_(momentum).add_(g)
        p.add_(buf, alpha=-lr)


def rmsprop_step(params, grads, st, lr=1e-3, alpha=0.9, eps=1e-8, wd=0.0):
    st.step += 1
    for p, g, sq in zip(params, grads, st.s1):
        g = g.add(p, alpha=wd)
        sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
        p.addcdiv_(g, sq.sqrt().add_(eps), value=-lr)


def adam_step(params, grads, st, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0):
    st.step += 1
    bc1 = 1 - beta1 ** st.step
    bc2 = 1 - beta2 ** st.step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(params, grads, st.s1, st.s2):
        g = g.add(p, alpha=wd)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


OPTIM_STEPS = {'sgd': sgd_step, 'rmsprop': rmsprop_step, 'adam': adam_step}
