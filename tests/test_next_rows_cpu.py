"""SURVEY.md 8(f) rows on the CPU: the oracle restatements of the non-AdamW optimizers against fixtures produced by
the reference's own get_optimizer (oracle/make_golden_optim.py -> tests/golden/optim_steps.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import optim as ooptim
from oracle.make_golden_optim import HYPER, STEPS, synth_problem


def _run_oracle(name, dtype=torch.float32):
    p0, grads = synth_problem()
    params = [p.clone().to(dtype) for p in p0]
    st = ooptim.AdamWState(params) if name == 'adamw' else ooptim.FlatState(params)
    traj = []
    for s in range(STEPS):
        gl = [g.clone().to(dtype) for g in grads[s]]
        ooptim.clip_grad_norm(gl, 1.0)
        if name == 'adamw':
            ooptim.adamw_step(params, gl, st, lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, wd=HYPER['weight_decay'])
        elif name == 'adam':
            ooptim.adam_step(params, gl, st, lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, wd=HYPER['weight_decay'])
        elif name == 'sgd':
            ooptim.sgd_step(params, gl, st, lr=HYPER['lr'], momentum=HYPER['momentum'], wd=HYPER['weight_decay'])
        else:
            ooptim.rmsprop_step(params, gl, st, lr=HYPER['lr'], alpha=HYPER['momentum'], eps=1e-8, wd=HYPER['weight_decay'])
        traj.append(torch.cat([p.reshape(-1) for p in params]).double().numpy().copy())
    return np.stack(traj)


@pytest.mark.parametrize('name', ['sgd', 'rmsprop', 'adam', 'adamw'])
def test_oracle_optimizers_match_reference_factory(golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'optim_steps.npz'))[name]
    got = _run_oracle(name)
    # same fp32 arithmetic as torch.optim: agreement to a few ulps of the parameter scale
    np.testing.assert_allclose(got, g, rtol=2e-6, atol=2e-7)
    # the trajectories are not trivially equal to the start (every step moved the parameters)
    p0 = torch.cat([p.reshape(-1) for p in synth_problem()[0]]).numpy()
    assert np.abs(g[0] - p0).max() > 1e-4 and np.abs(g[-1] - g[0]).max() > 1e-4
