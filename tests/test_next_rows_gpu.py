"""SURVEY.md 8(f) rows on the GPU, through the C-ABI: fused SGD / RMSprop / Adam over the flat buffer against the
oracle (float64) and against the fixtures the reference's own get_optimizer produced (tests/golden/optim_steps.npz)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from speedplusbaseline_b200 import _lib as L
from oracle import optim as ooptim
from oracle.make_golden_optim import HYPER, STEPS, synth_problem
from kutil import rel, sp

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _hp_block(**kw):
    h = L.AdamWHp()
    for k, v in kw.items():
        setattr(h, k, v)
    return torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).cuda()


KINDS = {'adamw': L.OPT_ADAMW, 'sgd': L.OPT_SGD, 'rmsprop': L.OPT_RMSPROP, 'adam': L.OPT_ADAM}


@pytest.mark.parametrize('name', ['sgd', 'rmsprop', 'adam', 'adamw'])
def test_optim_step_matches_reference_factory_golden(golden_dir, name):
    """raw C-ABI call on a 1031-element buffer (vector body + scalar tail), clip_grad_norm_(1.0) folded in."""
    gold = np.load(os.path.join(golden_dir, 'optim_steps.npz'))[name]
    p0, grads = synth_problem()
    p = torch.cat([t.reshape(-1) for t in p0]).cuda()
    s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
    hp = _hp_block(lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, weight_decay=HYPER['weight_decay'],
                   max_norm=1.0, clip_value=1.0, grad_scale=1.0, step=0, clip_mode=1)
    for s in range(STEPS):
        g = torch.cat([t.reshape(-1) for t in grads[s]]).cuda()
        L.call('b200sp_grad_sqnorm', g.data_ptr(), p.numel(), hp.data_ptr(), sp())
        L.call('b200sp_optim_step', KINDS[name], p.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, p.numel(),
               hp.data_ptr(), sp())
        torch.cuda.synchronize()
        np.testing.assert_allclose(p.cpu().numpy(), gold[s], rtol=3e-6, atol=3e-7, err_msg='%s step %d' % (name, s))
    h = L.AdamWHp.from_buffer_copy(hp.cpu().numpy().tobytes())
    assert h.step == STEPS and h.sqnorm == 0.0


@pytest.mark.parametrize('name,clip_mode', [('sgd', 0), ('sgd', 1), ('rmsprop', 1), ('rmsprop', 2), ('adam', 0), ('adam', 2)])
def test_fused_optimizer_classes_match_oracle(name, clip_mode):
    from speedplusbaseline_b200.params import ParamStore
    from speedplusbaseline_b200 import optim as fo
    g = _g(31 + clip_mode + len(name))
    st = ParamStore([('a.weight', 'plain', (37, 5)), ('b.weight', 'plain', (1001,))], [('bn', 8)], torch.device('cuda'))
    p0 = torch.randn(st.n, generator=g)
    st.params.copy_(p0)
    par = [torch.nn.Parameter(st.params)]
    if name == 'sgd':
        opt = fo.FusedSGD(st, par, lr=1e-2, momentum=0.9, weight_decay=0.01, clip_mode=clip_mode)
    elif name == 'rmsprop':
        opt = fo.FusedRMSprop(st, par, lr=1e-3, alpha=0.9, weight_decay=0.01, clip_mode=clip_mode)
    else:
        opt = fo.FusedAdam(st, par, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=clip_mode)
    ref_p = [p0.clone().double()]
    ost = ooptim.FlatState(ref_p)
    for it in range(4):
        gr = torch.randn(st.n, generator=g) * (3.0 if it == 1 else 0.01)
        st.grads.copy_(gr)
        opt.step()
        gl = [gr.clone().double()]
        if clip_mode == 1:
            ooptim.clip_grad_norm(gl, 1.0)
        elif clip_mode == 2:
            ooptim.clip_grad_value(gl, 1.0)
        if name == 'sgd':
            ooptim.sgd_step(ref_p, gl, ost, lr=1e-2, momentum=0.9, wd=0.01)
        elif name == 'rmsprop':
            ooptim.rmsprop_step(ref_p, gl, ost, lr=1e-3, alpha=0.9, eps=1e-8, wd=0.01)
        else:
            ooptim.adam_step(ref_p, gl, ost, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01)
        torch.cuda.synchronize()
        assert rel(st.params, ref_p[0]) < 2e-6, it
    sd = opt.state_dict()
    assert sd['state']['step'] == 4 and (sd['state']['exp_avg_sq'] is None) == (name != 'adam')
    assert rel(opt.exp_avg, ost.s1[0]) < 1e-5


def test_optim_step_rejects_bad_arguments():
    p = torch.zeros(8, device='cuda')
    hp = _hp_block(lr=1e-3)
    assert L.lib.b200sp_optim_step(7, p.data_ptr(), p.data_ptr(), p.data_ptr(), None, None, 8, hp.data_ptr(), sp()) == -22
    assert L.lib.b200sp_optim_step(L.OPT_ADAM, p.data_ptr(), p.data_ptr(), p.data_ptr(), None, None, 8, hp.data_ptr(), sp()) == -22
