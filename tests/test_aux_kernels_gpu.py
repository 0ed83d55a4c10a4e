"""Kernel-level parity of the SPN / DANN / style-aug entry points against the matching torch ops (fp64 on CPU)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from speedplusbaseline_b200 import _lib as L
from kutil import rel, sp, vt_plain

pytestmark = pytest.mark.gpu
dev = 'cuda'


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('Cin,c_off,Cg,k,s,p,nchw', [(3, 0, 3, 11, 4, 0, 1), (96, 48, 48, 5, 1, 2, 0), (64, 0, 64, 3, 1, 1, 0), (32, 16, 16, 3, 2, 1, 0)])
def test_im2col_and_col2im(Cin, c_off, Cg, k, s, p, nchw):
    torch.manual_seed(0)
    B, H, W = 2, 23, 19
    x = torch.randn(B, Cin, H, W)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    K = k * k * Cg
    Kp = (K + 3) // 4 * 4
    xin = (x if nchw else nhwc(x)).to(dev)
    col = torch.full((B * Ho * Wo, Kp), 7.0, device=dev)
    L.call('b200sp_im2col', xin.data_ptr(), col.data_ptr(), B, H, W, Cin, c_off, Cg, k, s, p, Kp, nchw, sp())
    ref = F.unfold(x[:, c_off:c_off + Cg].double(), k, padding=p, stride=s)            # [B, Cg*k*k, L], row = c*k*k + kh*k + kw
    ref = ref.view(B, Cg, k * k, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, K)   # column = (kh*k+kw)*Cg + c
    assert torch.equal(col[:, :K].cpu().double(), ref)
    assert float(col[:, K:].abs().sum()) == 0
    if nchw:
        return
    # col2im = adjoint of im2col (+ ReLU mask)
    dcol = torch.randn(B * Ho * Wo, Kp, device=dev)
    act = torch.randn(B, H, W, Cin, device=dev)
    dx = torch.zeros(B, H, W, Cin, device=dev)
    L.call('b200sp_col2im', dcol.data_ptr(), dx.data_ptr(), act.data_ptr(), B, H, W, Cin, c_off, Cg, k, s, p, Kp, sp())
    d = dcol[:, :K].cpu().double().view(B, Ho * Wo, k * k, Cg).permute(0, 3, 2, 1).reshape(B, Cg * k * k, Ho * Wo)
    refx = F.fold(d, (H, W), k, padding=p, stride=s)                                    # [B, Cg, H, W]
    refx = refx * (act.cpu()[..., c_off:c_off + Cg].permute(0, 3, 1, 2) > 0)
    assert rel(dx[..., c_off:c_off + Cg].permute(0, 3, 1, 2), refx) < 1e-6


@pytest.mark.parametrize('lrn', [0, 1])
def test_pool_lrn_forward_backward(lrn):
    torch.manual_seed(1)
    B, H, W, Cc = 2, 13, 15, 24
    x = torch.relu(torch.randn(B, Cc, H, W)).double()            # post-ReLU input: plenty of ties at zero
    x.requires_grad_(True)
    p = F.max_pool2d(x, 3, 2)
    y = F.local_response_norm(p, 2, 2e-5, 0.75, 1.0) if lrn else p
    g = torch.randn_like(y)
    y.backward(g)
    Ho, Wo = p.shape[2], p.shape[3]
    xd = nhwc(x.detach().float()).to(dev)
    pooled, out = torch.zeros(B, Ho, Wo, Cc, device=dev), torch.zeros(B, Ho, Wo, Cc, device=dev)
    amax = torch.zeros(B, Ho, Wo, Cc, dtype=torch.uint8, device=dev)
    L.call('b200sp_pool_lrn_fwd', xd.data_ptr(), pooled.data_ptr(), out.data_ptr(), amax.data_ptr(), B, H, W, Cc, lrn, 2e-5, 0.75, sp())
    assert rel(out.permute(0, 3, 1, 2), y.detach()) < 1e-6
    gd = nhwc(g.float()).to(dev)
    scratch, dx = torch.zeros_like(pooled), torch.zeros_like(xd)
    L.call('b200sp_pool_lrn_bwd', gd.data_ptr(), pooled.data_ptr(), xd.data_ptr(), amax.data_ptr(), scratch.data_ptr(), dx.data_ptr(),
           B, H, W, Cc, lrn, 2e-5, 0.75, 1, sp())
    ref = x.grad * (x.detach() > 0)
    assert rel(dx.permute(0, 3, 1, 2), ref) < 1e-5


def test_soft_ce_and_dropout():
    torch.manual_seed(2)
    B, N = 5, 5000
    z = torch.randn(B, N) * 3
    t = torch.zeros(B, N)
    t[:, :5] = 0.2
    zd = z.double().requires_grad_(True)
    loss = (-(t.double() * F.log_softmax(zd, 1)).sum(1)).mean()
    (10.0 * loss).backward()
    rows, dz = torch.zeros(B, device=dev), torch.zeros(B, N, device=dev)
    zg, tg = z.to(dev), t.to(dev)                      # keep the device copies alive across the launch
    L.call('b200sp_soft_ce', zg.data_ptr(), tg.data_ptr(), rows.data_ptr(), dz.data_ptr(), B, N, 10.0, sp())
    assert float(rows.mean()) == pytest.approx(float(loss.detach()), rel=1e-5)
    assert rel(dz, zd.grad) < 1e-5
    x = torch.randn(1 << 20, device=dev)
    out, mask = torch.empty_like(x), torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    L.call('b200sp_dropout_fwd', x.data_ptr(), out.data_ptr(), mask.data_ptr(), x.numel(), 0.5, 12345, sp())
    assert 0.49 < float(mask.float().mean()) < 0.51
    assert torch.equal(out, x * mask.float() * 2)
    L.call('b200sp_dropout_fwd', x.data_ptr(), out.data_ptr(), mask.data_ptr(), x.numel(), 0.5, 12346, sp())
    m2 = mask.clone()
    L.call('b200sp_dropout_fwd', x.data_ptr(), out.data_ptr(), mask.data_ptr(), x.numel(), 0.5, 12345, sp())
    assert not torch.equal(m2, mask)                                  # the mask depends on the seed


def test_strided_group_gemms_and_splitk_fc():
    torch.manual_seed(3)
    M, O, g, Kg = 300, 64, 2, 40
    Og = O // g
    col = [torch.randn(M, Kg, device=dev) for _ in range(g)]
    w = torch.randn(O, Kg, device=dev) * 0.1
    bias = torch.randn(O, device=dev)
    y = torch.zeros(M, O, device=dev)
    for gi in range(g):
        v = vt_plain(col[gi])
        L.call('b200sp_gemm_fwd', C.byref(v), Kg, w.data_ptr() + 4 * gi * Og * Kg, bias.data_ptr() + 4 * gi * Og, L.ACT_RELU,
               y.data_ptr() + 4 * gi * Og, O, M, Og, Kg, L.F32, sp())
    ref = torch.cat([torch.relu(col[gi].double() @ w[gi * Og:(gi + 1) * Og].double().t() + bias[gi * Og:(gi + 1) * Og].double()) for gi in range(g)], 1)
    assert rel(y, ref) < 1e-5
    dy = torch.randn(M, O, device=dev)
    for gi in range(g):
        vdy, vc = L.VTensor(dy.data_ptr() + 4 * gi * Og, None, None, None, None, 0, 0), vt_plain(col[gi])
        dw = torch.zeros(Og, Kg, device=dev)
        L.call('b200sp_gemm_wgrad', C.byref(vdy), O, C.byref(vc), Kg, dw.data_ptr(), M, Og, Kg, L.F32, sp())
        assert rel(dw, dy[:, gi * Og:(gi + 1) * Og].double().t() @ col[gi].double()) < 1e-5
        dcol = torch.zeros(M, Kg, device=dev)
        L.call('b200sp_gemm_dgrad', C.byref(vdy), O, w.data_ptr() + 4 * gi * Og * Kg, None, 1.0, dcol.data_ptr(), None, M, Og, Kg, L.F32, sp())
        assert rel(dcol, dy[:, gi * Og:(gi + 1) * Og].double() @ w[gi * Og:(gi + 1) * Og].double()) < 1e-5
    # split-K FC (one M tile)
    B, N, K = 32, 520, 4096
    x, wf, bf = torch.randn(B, K, device=dev), torch.randn(N, K, device=dev) * 0.02, torch.randn(N, device=dev)
    yf = torch.zeros(B, N, device=dev)
    L.call('b200sp_fc_fwd_splitk', x.data_ptr(), wf.data_ptr(), yf.data_ptr(), B, N, K, sp())
    L.call('b200sp_bias_act', yf.data_ptr(), bf.data_ptr(), B, N, 1, sp())
    assert rel(yf, torch.relu(x.double() @ wf.double().t() + bf.double())) < 1e-5
    dyf = torch.randn(B, N, device=dev)
    dx = torch.ones(B, K, device=dev)                                     # accumulates onto what is there
    L.call('b200sp_fc_dgrad_splitk', dyf.data_ptr(), wf.data_ptr(), dx.data_ptr(), B, N, K, sp())
    assert rel(dx, 1.0 + dyf.double() @ wf.double()) < 1e-5


def test_dann_head_and_bce():
    torch.manual_seed(4)
    B, HW, Cc = 5, 49, 1280
    h = torch.relu(torch.randn(B, HW, Cc))
    w3, b3 = torch.randn(Cc) * 0.05, torch.randn(1)
    hd = h.double().requires_grad_(True)
    w3d, b3d = w3.double().requires_grad_(True), b3.double().requires_grad_(True)
    z = hd.mean(1) @ w3d + b3d
    loss = F.binary_cross_entropy_with_logits(z, torch.ones(B, dtype=torch.float64))
    loss.backward()
    hg, w3g, b3g = h.to(dev), w3.to(dev), b3.to(dev)
    pooled, zz = torch.zeros(B, Cc, device=dev), torch.zeros(B, device=dev)
    L.call('b200sp_dann_head_fwd', hg.data_ptr(), w3g.data_ptr(), b3g.data_ptr(), pooled.data_ptr(), zz.data_ptr(), B, HW, Cc, L.F32, sp())
    assert rel(zz, z.detach()) < 1e-5
    l, dz = torch.zeros(1, device=dev), torch.zeros(B, device=dev)
    L.call('b200sp_bce_logits', zz.data_ptr(), 1.0, l.data_ptr(), dz.data_ptr(), None, B, sp())
    assert float(l) == pytest.approx(float(loss.detach()), rel=1e-5)
    dw3, db3 = torch.zeros(Cc, device=dev), torch.zeros(1, device=dev)
    L.call('b200sp_dann_head_bwd', hg.data_ptr(), dz.data_ptr(), pooled.data_ptr(), w3g.data_ptr(), dw3.data_ptr(), db3.data_ptr(),
           B, HW, Cc, L.F32, sp())
    assert rel(dw3, w3d.grad) < 1e-5 and rel(db3, b3d.grad) < 1e-5
    assert rel(hg, hd.grad * (h.double() > 0)) < 1e-5                 # in place: dL/d(pre-ReLU conv output)
    s = torch.tensor([-0.37], device=dev)
    before = hg.clone()
    L.call('b200sp_scale_dev', hg.data_ptr(), hg.numel(), s.data_ptr(), 1.0, L.F32, sp())
    assert torch.allclose(hg, before * -0.37)


@pytest.mark.parametrize('pad,up,ps', [(1, 1, 1), (4, 1, 1), (1, 2, 1), (1, 1, 2)])
def test_in_apply_writes_the_next_conv_planes(pad, up, ps):
    """normalise + affine + ReLU then reflection pad / x2 nearest upsample / stride-2 phase split == torch on the result"""
    torch.manual_seed(5)
    B, Hs, Ws, Cc = 2, 12, 16, 32
    raw = torch.randn(B, Hs, Ws, Cc)
    scale, shift = torch.rand(B, Cc) + 0.5, torch.randn(B, Cc) * 0.2
    v = torch.relu(raw * scale[:, None, None, :] + shift[:, None, None, :]).permute(0, 3, 1, 2)     # NCHW
    if up == 2:
        v = F.interpolate(v, scale_factor=2)
    ref = F.pad(v, (pad,) * 4, mode='reflect')                                                       # [B,C,Hp,Wp]
    Hd, Wd = ref.shape[2] // ps, ref.shape[3] // ps
    planes = [torch.zeros(B * Hd * Wd + 16, Cc, dtype=torch.bfloat16, device=dev) for _ in range(ps * ps)]
    d = L.InApplyDesc()
    rd, sd_, hd_ = raw.to(dev), scale.to(dev).contiguous(), shift.to(dev).contiguous()
    d.raw, d.scale, d.shift, d.res_in, d.res_out = rd.data_ptr(), sd_.data_ptr(), hd_.data_ptr(), None, None
    for q in range(4):
        d.planes[q] = planes[q].data_ptr() if q < len(planes) else None
    d.B, d.Hs, d.Ws, d.Cs, d.C, d.act, d.pad, d.up, d.ps, d.Hd, d.Wd, d.Cd = B, Hs, Ws, Cc, Cc, L.ACT_RELU, pad, up, ps, Hd, Wd, Cc
    L.call('b200sp_in_apply', C.byref(d), sp())
    for qy in range(ps):
        for qx in range(ps):
            got = planes[qy * ps + qx][:B * Hd * Wd].view(B, Hd, Wd, Cc).permute(0, 3, 1, 2).float().cpu()
            want = ref[:, :, qy::ps, qx::ps].bfloat16().float()
            assert torch.equal(got, want), (qy, qx)
