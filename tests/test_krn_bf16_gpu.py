"""--use_fp16 path: bf16 activation storage + bf16 tensor-core GEMMs (fp32 accumulate, fp32 BatchNorm statistics,
fp32 master weights / gradients / optimizer).  The reference's own mixed-precision mode (torch autocast fp16) is
not bit-comparable either; the bar here is closeness to the float64 oracle at bf16 resolution: eval logits within
5e-2 relative L2, train loss within 5e-2, gradients aligned (cosine > 0.97 on the large tensors), and a few AdamW
steps on a fixed batch must reduce the loss."""
import pytest
import torch

from oracle import krn as okrn, synth, steps
from kutil import rel

pytestmark = pytest.mark.gpu


def _model(sd):
    from speedplusbaseline_b200 import _lib as L
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    m = KeypointRegressionNet(11, device='cuda:0', dtype=L.BF16)
    m.load_state_dict(sd)
    return m


def test_eval_logits_bf16():
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    x = synth.synth_images(4, seed=5)
    with torch.no_grad():
        xr, yr = okrn.krn_forward({k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}, x.double())
    xc, yc = _model(sd).eval()(x.cuda())
    e = rel(torch.cat([xc, yc], 1), torch.cat([xr, yr], 1))
    assert e < 5e-2, e


def test_train_step_bf16_close_to_oracle_and_learns():
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    B = 8
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    x, y = synth.synth_images(B), synth.synth_keypoints(B)
    s64 = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    r64 = steps.krn_train_step(s64, steps.new_state(s64), x.double(), y.double())
    m = _model(sd).train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=1)
    stp = KRNTrainStep(m, opt, use_graph=False)
    cx = stp._fwd_bwd(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    assert float(cx.loss3[0]) == pytest.approx(r64['loss'], rel=5e-2)
    gd = m.grad_dict()
    for k in ('head.0.weight', 'extras.3.conv.3.weight', 'extras.1.conv.3.weight', 'base.17.conv.2.weight', 'base.8.conv.0.0.weight',
              'base.2.conv.0.0.weight', 'base.1.conv.0.0.weight'):
        a, b = gd[k].double().cpu().flatten(), r64['grads'][k].flatten()
        cos = float((a @ b) / (a.norm() * b.norm()))
        assert cos > 0.97, (k, cos)
        assert 0.8 < float(a.norm() / b.norm()) < 1.25, (k, float(a.norm() / b.norm()))
    losses = []
    xg, yg = x.cuda(), y.cuda()
    for _ in range(6):
        losses.append(float(stp.step(xg, yg)[0]))
    assert all(l == l and l < 1e6 for l in losses), losses
    assert losses[-1] < losses[0], losses
    # the bf16 weight mirror follows the fp32 master
    st = m._store
    assert torch.equal(st.params_lowp.float(), st.params.bfloat16().float())
