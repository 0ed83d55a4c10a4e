"""Generate tests/golden/optim_steps.npz from the UNMODIFIED reference factory
(/root/reference/src/nets/build.py:60-78 get_optimizer -> torch.optim.{SGD,RMSprop,Adam,AdamW}) driven the way
the reference loop drives it (zero_grad, backward, clip_grad_norm_(params, 1.0), step; trainer.py:87-98).

Build container only (the reference does not travel to the GPU box):  python -m oracle.make_golden_optim

A tiny two-tensor module (1031 parameters: exercises the vector body AND the scalar tail of the flat kernels), seeded
gradients scaled so that the clip is active on some steps and inactive on others, 4 steps per optimizer.
"""
import os
import sys
import types

import numpy as np

from oracle.make_golden import OUT, _shims

SHAPES = [(32, 32), (7,)]
STEPS = 4
HYPER = dict(lr=1e-2, momentum=0.9, weight_decay=0.01)


def synth_problem():
    import torch
    g = torch.Generator().manual_seed(77)
    params = [torch.randn(s, generator=g) * 0.5 for s in SHAPES]
    # gradient scale per step: norm > 1 (clip active) on steps 0 and 2, < 1 on steps 1 and 3
    grads = [[torch.randn(s, generator=g) * sc for s in SHAPES] for sc in (0.2, 0.01, 1.0, 0.005)]
    return params, grads


def main():
    _shims()
    import torch
    from config import cfg
    from src.nets.build import get_optimizer
    from torch.nn.utils import clip_grad_norm_
    out = {}
    for name in ('sgd', 'rmsprop', 'adam', 'adamw'):
        p0, grads = synth_problem()
        m = torch.nn.Module()
        m.a, m.b = torch.nn.Parameter(p0[0].clone()), torch.nn.Parameter(p0[1].clone())
        cfg.optimizer, cfg.lr, cfg.momentum, cfg.weight_decay = name, HYPER['lr'], HYPER['momentum'], HYPER['weight_decay']
        opt = get_optimizer(cfg, m)
        traj = []
        for s in range(STEPS):
            opt.zero_grad(set_to_none=True)
            m.a.grad, m.b.grad = grads[s][0].clone(), grads[s][1].clone()
            clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            traj.append(torch.cat([m.a.detach().reshape(-1), m.b.detach().reshape(-1)]).numpy().copy())
        out[name] = np.stack(traj)
    np.savez(os.path.join(OUT, 'optim_steps.npz'), **out)
    print('written', os.path.join(OUT, 'optim_steps.npz'), {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
