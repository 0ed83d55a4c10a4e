"""Optimizer checkpoints in the reference's format (SURVEY.md 5 "Boundary contract"; reference utils.py:109-135 saves
`optimizer.state_dict()` of torch.optim.AdamW): per-parameter {step, exp_avg, exp_avg_sq} in the reference shapes, indexed in
the reference module's parameters() order.  Both directions, plus the device-resident step counter under CUDA graphs."""
import os

import pytest
import torch

from oracle import krn as okrn, synth
from kutil import rel

pytestmark = pytest.mark.gpu


def _torch_reference_optimizer(sd, steps, seed=0):
    """What the reference's get_optimizer builds (build.py:72-74): torch.optim.AdamW over model.parameters(), i.e. the
    state_dict order without the BatchNorm buffers."""
    keys = [k for k in sd if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]
    params = [torch.nn.Parameter(sd[k].clone()) for k in keys]
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    g = torch.Generator().manual_seed(seed)
    grads = []
    for _ in range(steps):
        gs = [torch.randn(p.shape, generator=g) * 1e-2 for p in params]
        for p, gr in zip(params, gs):
            p.grad = gr
        opt.step()
        grads.append(gs)
    return keys, params, opt, grads


def _model():
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.optim import FusedAdamW
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    m = KeypointRegressionNet(11, device='cuda:0', seed=1)
    m.load_state_dict(sd)
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=0)
    return sd, m, opt


def test_resume_from_a_torch_adamw_checkpoint_and_write_one_back():
    sd, m, opt = _model()
    keys, params, topt, _ = _torch_reference_optimizer(sd, steps=3)
    ck = topt.state_dict()                                   # the dict the reference stores under 'optimizer'
    m.load_state_dict({**sd, **{k: p.detach() for k, p in zip(keys, params)}})
    opt.load_state_dict(ck)
    assert opt.device_step() == 3
    out = opt.state_dict()
    assert set(out) == {'state', 'param_groups'} and len(out['state']) == len(keys) == len(out['param_groups'][0]['params'])
    for i in (0, 1, 5, len(keys) // 2, len(keys) - 2, len(keys) - 1):
        for f in ('exp_avg', 'exp_avg_sq'):
            assert out['state'][i][f].shape == ck['state'][i][f].shape
            assert torch.equal(out['state'][i][f].cpu(), ck['state'][i][f]), (i, f)
        assert float(out['state'][i]['step']) == 3.0
    # a torch optimizer (= the reference, or this repo's --no_cuda path) accepts what we write
    topt2 = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in params], lr=1e-3, weight_decay=0.01)
    topt2.load_state_dict(out)
    # one more identical step on both sides: parameters agree (bias correction uses step 4 on both)
    g = torch.Generator().manual_seed(9)
    gs = [torch.randn(p.shape, generator=g) * 1e-2 for p in params]
    for p, gr in zip(topt2.param_groups[0]['params'], gs):
        p.grad = gr
    topt2.step()
    gd = dict(zip(keys, gs))
    from speedplusbaseline_b200.params import to_native
    st = m._store
    for k, e in st.entries.items():
        st.grads[e.off:e.off + e.numel].copy_(to_native(e.kind, gd[k]).reshape(-1))
    for i, (pfx, _) in enumerate(st.bns):
        gsl, bsl, _, _ = st.bn_slices(i)
        st.grads[gsl].copy_(gd[pfx + '.weight'])
        st.grads[bsl].copy_(gd[pfx + '.bias'])
    opt.step()
    torch.cuda.synchronize()
    new = m.state_dict()
    worst = max(rel(new[k], p) for k, p in zip(keys, topt2.param_groups[0]['params']))
    assert worst < 1e-6, worst


def test_step_counter_survives_cuda_graph_replays():
    """graph replays advance only the device counter; state_dict() must read it, and a reload must restore it."""
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    sd, m, opt = _model()
    m.train()
    stepper = KRNTrainStep(m, opt, use_graph=True)
    x, y = synth.synth_images(2).cuda(), synth.synth_keypoints(2).cuda()
    for _ in range(5):
        stepper.step(x, y)
    torch.cuda.synchronize()
    ck = opt.state_dict()
    assert float(ck['state'][0]['step']) == 5.0
    sd2, m2, opt2 = _model()
    opt2.load_state_dict(ck)
    assert opt2.device_step() == 5
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)


def test_round_trip_through_the_reference_checkpoint_helpers(tmp_path):
    """save with the reference's own save_checkpoint / load with its load_checkpoint (utils.py:109-135) when a staged copy exists."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline')
    if not os.path.exists(os.path.join(root, '_ref', 'src', 'utils', 'utils.py')):
        pytest.skip('reference not staged')
    import sys
    sys.path.insert(0, root)
    import ref_runner
    ref_runner._import_reference()
    from src.utils.utils import save_checkpoint, load_checkpoint
    sd, m, opt = _model()
    keys, params, topt, _ = _torch_reference_optimizer(sd, steps=2)
    save_checkpoint({'epoch': 7, 'model': 'krn', 'state_dict': sd, 'best_score': 7, 'optimizer': topt.state_dict()}, False, str(tmp_path))
    last_epoch, _ = load_checkpoint(os.path.join(str(tmp_path), 'checkpoint.pth.tar'), m, opt, torch.device('cuda:0'))
    assert last_epoch == 7 and opt.device_step() == 2
    got = opt.state_dict()
    assert torch.equal(got['state'][3]['exp_avg'].cpu(), topt.state_dict()['state'][3]['exp_avg'])
