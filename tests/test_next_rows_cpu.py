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


# ---- row f2: evaluation tail ----------------------------------------------------------------------
def _tie_free_rows(w, k):
    """rows whose k+1 largest values are pairwise distinct (torch.topk's order among exact ties is unspecified)."""
    s = -np.sort(-w, axis=1)[:, :k + 1]
    return np.array([len(np.unique(r)) == k + 1 for r in s])


def test_oracle_postproc_matches_reference_loops(golden_dir):
    from oracle import postproc
    from oracle.make_golden_postproc import K_NB, synth_inputs
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    w, x, y, bb = (t.numpy() for t in synth_inputs())
    tw, ti = postproc.spn_top_classes(w, K_NB)
    free = _tie_free_rows(w, K_NB)
    assert free.sum() >= 4 and not free.all()              # the fixture holds both kinds of rows
    assert (ti[free] == g['top_idx'][free]).all()          # bit-exact indices where the order is defined
    # every row, ties included: the selected VALUES are the reference's, in the same (descending) order
    assert (np.take_along_axis(w, ti, 1) == np.take_along_axis(w, g['top_idx'], 1)).all()
    np.testing.assert_allclose(tw, g['top_w'], rtol=1e-6, atol=1e-9)
    assert (postproc.krn_keypoints_pix(x, y, bb) == g['kpt_pix']).all()       # fp32, bit-exact
