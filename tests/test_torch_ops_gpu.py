"""`torch.ops.b200sp.*` (speedplusbaseline_b200/torch_ops.py): the libb200sp kernels registered with the torch dispatcher,
checked against the torch ops they replace (float64 on the CPU) through the dispatcher entry points."""
import pytest
import torch
import torch.nn.functional as F

from kutil import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    import speedplusbaseline_b200.torch_ops as T
    for n in T.OPS:
        assert hasattr(torch.ops.b200sp, n)
    return torch.ops.b200sp


def test_conv1x1_family(ops):
    g = torch.Generator().manual_seed(0)
    B, H, W, K, N = 3, 14, 14, 96, 64
    x = torch.randn(B, H, W, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.2
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    xin = torch.clamp(x.double() * sc.double() + sh.double(), 0, 6)
    ref = xin.reshape(-1, K) @ w.double().t()
    y, mean, rstd, scale, shift = ops.conv1x1_fwd(x.cuda(), w.cuda(), sc.cuda(), sh.cuda(), 2, gamma.cuda(), beta.cuda(), 1e-5)
    assert y.shape == (B, H, W, N) and rel(y.reshape(-1, N), ref) < 2e-5
    assert rel(mean, ref.mean(0)) < 1e-5 and rel(rstd, 1 / torch.sqrt(ref.var(0, unbiased=False) + 1e-5)) < 1e-5
    assert rel(scale, gamma.double() / torch.sqrt(ref.var(0, unbiased=False) + 1e-5)) < 1e-5
    dy = torch.randn(B, H, W, N, generator=g)
    dx = ops.conv1x1_dgrad(dy.cuda(), w.cuda())
    assert rel(dx.reshape(-1, K), dy.double().reshape(-1, N) @ w.double()) < 2e-5
    dw = ops.conv1x1_wgrad(dy.cuda(), x.cuda(), sc.cuda(), sh.cuda(), 2)
    assert rel(dw, dy.double().reshape(-1, N).t() @ xin.reshape(-1, K)) < 5e-5
    y2 = ops.conv1x1_fwd(x.cuda(), w.cuda(), None, None, 0, None, None, 0.0)[0]
    assert rel(y2.reshape(-1, N), x.double().reshape(-1, K) @ w.double().t()) < 2e-5


def test_depthwise_bn_reorg_loss(ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 14, 14, 32, generator=g)
    w = torch.randn(32, 1, 3, 3, generator=g)
    for s in (1, 2):
        y = ops.conv_dw3x3_fwd(x.cuda(), w.cuda(), s, None, None, 0)
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, s, 1, 1, 32).permute(0, 2, 3, 1)
        assert rel(y, ref) < 1e-6
    sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g)
    res = torch.randn(2, 14, 14, 32, generator=g)
    out = ops.bn_apply(x.cuda(), sc.cuda(), sh.cuda(), res.cuda(), 1)
    assert rel(out, torch.relu(x.double() * sc.double() + sh.double()) + res.double()) < 1e-6     # activation, THEN the residual (mobilenetv2.py:61-62)
    xr, x1 = torch.randn(2, 14, 14, 8, generator=g), torch.randn(2, 7, 7, 16, generator=g)
    cat = ops.reorg_cat(xr.cuda(), x1.cuda())
    t = xr.permute(0, 3, 1, 2)                                   # park2019.py:70-80 on NCHW
    B, Cc, Hh, Ww = t.shape
    t = t.view(B, Cc, Hh // 2, 2, Ww // 2, 2).transpose(3, 4).contiguous().view(B, Cc, Hh // 2 * Ww // 2, 4).transpose(2, 3).contiguous()
    t = t.view(B, Cc, 4, Hh // 2, Ww // 2).transpose(1, 2).contiguous().view(B, 4 * Cc, Hh // 2, Ww // 2)
    refc = torch.cat((t, x1.permute(0, 3, 1, 2)), 1).permute(0, 2, 3, 1)
    assert torch.equal(cat.cpu(), refc)
    logits, target = torch.randn(6, 22, generator=g), torch.rand(6, 2, 11, generator=g)
    loss3, dl = ops.krn_loss(logits.cuda(), target.cuda())
    lg = logits.double().clone().requires_grad_(True)
    lx = sum(F.mse_loss(lg[:, 2 * i], target[:, 0, i].double()) for i in range(11))
    ly = sum(F.mse_loss(lg[:, 2 * i + 1], target[:, 1, i].double()) for i in range(11))
    (lx + ly).backward()
    assert abs(float(loss3[0]) - float(lx + ly)) < 1e-5 * float(lx + ly) and rel(dl, lg.grad) < 1e-6


def test_adamw_fused_matches_torch(ops):
    g = torch.Generator().manual_seed(2)
    n = 100003
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.05
    p = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    pc, m, v = p0.cuda(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    for step in range(3):
        p.grad = gr.double().clone()
        total = torch.nn.utils.clip_grad_norm_([p], 1.0)
        opt.step()
        norm = ops.adamw_fused(pc, gr.cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 1.0)
        assert abs(float(norm) - float(total)) < 1e-5 * float(total)
        assert rel(pc, p.detach()) < 1e-6


def test_cpu_tensors_are_rejected(ops):
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.conv1x1_dgrad(torch.randn(4, 8), torch.randn(8, 8))
