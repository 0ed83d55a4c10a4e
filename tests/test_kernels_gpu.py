"""Kernel-level parity (through the C-ABI) against torch CPU float64 ops on identical inputs.
Covers ragged sizes (M not a tile multiple, narrow N/K), both strides, all activation codes."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from speedplusbaseline_b200 import _lib as L
from oracle import optim as ooptim
from kutil import BnB, BnF, act_t, rel, sp, vt_bnact, vt_dy, vt_plain

pytestmark = pytest.mark.gpu
TOL = 2e-5       # fp32 storage, 3xTF32 GEMMs / fp32 FMA stencils vs float64 truth


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize('M,N,K', [(301, 16, 32), (1000, 96, 16), (257, 24, 144), (2352, 320, 960), (64, 1024, 320), (5, 64, 96),
                                   (40003, 96, 16), (2352, 1024, 1280), (20001, 144, 24), (9408, 576, 96),
                                   # the long-M / tiny-N*K shapes of the direct fp32 kernels (pwdirect.cu), ragged M
                                   (9409, 16, 32), (12001, 24, 96), (9999, 24, 144), (10001, 32, 144), (9408, 192, 32), (11111, 32, 192),
                                   # BASELINE.json's full sizes (bs=48, 112x112 maps): M = 602,112
                                   (602112, 96, 16), (602112, 16, 32)])
@pytest.mark.parametrize('act', [L.ACT_NONE, L.ACT_RELU6])
def test_pw_fwd_with_bn_epilogue(M, N, K, act):
    g = _g(M + N + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.3
    xin = act_t(x.double() * sc.double() + sh.double(), act) if act else x.double()
    ref = xin @ w.double().t()
    xd, wd = x.cuda(), w.cuda()
    y = torch.empty(M, N, device='cuda')
    bn = BnF(N, gamma=(torch.rand(N, generator=g) + 0.5).cuda(), beta=torch.randn(N, generator=g).cuda())
    vt = vt_bnact(xd, sc.cuda(), sh.cuda(), act) if act else vt_plain(xd)
    L.call('b200sp_pw_fwd', C.byref(vt), wd.data_ptr(), None, 0, y.data_ptr(), bn.ref(), M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(y, ref) < TOL
    mean, var = ref.mean(0), ref.var(0, unbiased=False)
    assert rel(bn.mean, mean) < 1e-5 and rel(bn.rstd, 1 / torch.sqrt(var + 1e-5)) < 1e-5
    scale = bn.gamma.cpu().double() / torch.sqrt(var + 1e-5)
    assert rel(bn.scale, scale) < 1e-5 and rel(bn.shift, bn.beta.cpu().double() - mean * scale) < 1e-4
    assert rel(bn.rm, 0.1 * mean) < 1e-5
    assert rel(bn.rv, 0.9 + 0.1 * ref.var(0, unbiased=True)) < 1e-5
    assert float(bn.sum.abs().max()) == 0.0 and int(bn.ticket[0]) == 0      # workspace left clean


def test_pw_fwd_bias_act_no_bn():
    g = _g(1)
    M, N, K = 98, 1280, 320
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / 18, torch.randn(N, generator=g)
    y = torch.empty(M, N, device='cuda')
    xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
    L.call('b200sp_pw_fwd', C.byref(vt_plain(xc)), wc.data_ptr(), bc.data_ptr(), L.ACT_RELU, y.data_ptr(),
           None, M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(y, torch.relu(x.double() @ w.double().t() + b.double())) < TOL


@pytest.mark.parametrize('M,N,K', [(301, 16, 32), (999, 96, 16), (2352, 1024, 320), (130, 24, 144), (40003, 96, 16), (20001, 24, 144),
                                   (2352, 1024, 1280), (9408, 96, 576), (784, 64, 96),
                                   (9409, 16, 32), (12001, 24, 96), (20001, 144, 24), (10001, 32, 144), (9408, 192, 32), (11111, 32, 192),
                                   (602112, 96, 16)])
@pytest.mark.parametrize('act', [L.ACT_NONE, L.ACT_RELU6, L.ACT_LEAKY02])
def test_pw_dgrad_fused_bn_backward(M, N, K, act):
    g = _g(M * 3 + N + K + act)
    gq, yq = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    cA, cB, cC = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    w = torch.randn(N, K, generator=g) / N ** 0.5
    skip = torch.randn(M, K, generator=g)
    yprev = torch.randn(M, K, generator=g) * 2
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g)
    mean, rstd = torch.randn(K, generator=g) * 0.2, torch.rand(K, generator=g) + 0.5
    dy = cA.double() * gq.double() + cB.double() * yq.double() + cC.double()
    dx = dy @ w.double() + skip.double()
    z = yprev.double() * sc.double() + sh.double()
    zt = z.clone().requires_grad_(True)
    act_t(zt, act).sum().backward()
    gref = dx * zt.grad
    xhat = (yprev.double() - mean.double()) * rstd.double()
    s1, s2 = gref.sum(0), (gref * xhat).sum(0)
    d = {k: v.cuda() for k, v in dict(gq=gq, yq=yq, cA=cA, cB=cB, cC=cC, w=w, skip=skip, yprev=yprev, sc=sc, sh=sh,
                                      mean=mean, rstd=rstd).items()}
    out = torch.empty(M, K, device='cuda')
    bn = BnB(d['yprev'], d['sc'], d['sh'], d['mean'], d['rstd'], act)
    bn.dgamma.fill_(1.0)          # must accumulate, not overwrite
    L.call('b200sp_pw_dgrad', C.byref(vt_dy(d['gq'], d['yq'], d['cA'], d['cB'], d['cC'])), d['w'].data_ptr(),
           d['skip'].data_ptr(), 1.0, out.data_ptr(), bn.ref(), M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(out, gref) < TOL
    assert rel(bn.dbeta, s1) < 1e-5 and rel(bn.dgamma - 1.0, s2) < 1e-4
    cBr = -sc.double() * rstd.double() * s2 / M
    assert rel(bn.cA, sc) < 1e-6 and rel(bn.cB, cBr) < 1e-4
    assert rel(bn.cC, -sc.double() * s1 / M - cBr * mean.double()) < 1e-4
    assert float(bn.s1.abs().max()) == 0.0


def test_pw_dgrad_plain_scale():
    g = _g(9)
    M, N, K = 98, 1280, 320
    dy, w = torch.randn(M, N, generator=g), torch.randn(N, K, generator=g) / 30
    out = torch.empty(M, K, device='cuda')
    dyc, wc = dy.cuda(), w.cuda()
    L.call('b200sp_pw_dgrad', C.byref(vt_plain(dyc)), wc.data_ptr(), None, -0.37, out.data_ptr(), None,
           M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(out, -0.37 * (dy.double() @ w.double())) < TOL


@pytest.mark.parametrize('M,N,K', [(3001, 16, 32), (1999, 96, 16), (2352, 1024, 320), (4097, 24, 144), (100, 64, 96), (60003, 96, 16),
                                   (2352, 1024, 1280), (9408, 576, 96), (37632, 32, 192), (784, 64, 96), (196, 64, 96), (602112, 96, 16)])
def test_pw_wgrad(M, N, K):
    g = _g(M + 7 * N + K)
    gq, yq = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    cA, cB, cC = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    x = torch.randn(M, K, generator=g)
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g)
    dy = cA.double() * gq.double() + cB.double() * yq.double() + cC.double()
    xin = torch.clamp(x.double() * sc.double() + sh.double(), 0, 6)
    ref = dy.t() @ xin
    dw = torch.full((N, K), 0.5, device='cuda')
    db = torch.zeros(N, device='cuda')
    t = [v.cuda() for v in (gq, yq, cA, cB, cC, x, sc, sh)]
    L.call('b200sp_pw_wgrad', C.byref(vt_dy(*t[:5])), C.byref(vt_bnact(t[5], t[6], t[7], L.ACT_RELU6)), dw.data_ptr(),
           db.data_ptr(), M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(dw - 0.5, ref) < 5e-5
    assert rel(db, dy.sum(0)) < 1e-5


@pytest.mark.parametrize('M,N,K', [(70001, 16, 32), (65537, 24, 144), (80004, 96, 24), (66001, 192, 32), (65536, 32, 192), (70000, 24, 96),
                                   (602112, 16, 32), (150528, 24, 144)])
@pytest.mark.parametrize('dy_mode,x_mode', [('dy', 'bnact'), ('plain', 'plain'), ('dy', 'plain'), ('plain', 'bnact')])
def test_pw_wgrad_streaming_reduction(M, N, K, dy_mode, x_mode):
    """wgdirect.cu: the long-M / tiny-N*K weight gradients as an fp32 streaming reduction (M >= 65536, N*K <= 6144, N <= 4 K;
    the (96, 24) case stays on the tensor-core kernel): every operand-mode combination, ragged M (last slab partial),
    accumulation into a non-zero dW, one launch per call."""
    if M > 200000 and (dy_mode, x_mode) != ('dy', 'bnact'):
        pytest.skip('full size once')
    g = _g(M + 3 * N + K)
    gq, yq = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    cA, cB, cC = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    x = torch.randn(M, K, generator=g)
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g)
    dy = cA.double() * gq.double() + cB.double() * yq.double() + cC.double() if dy_mode == 'dy' else gq.double()
    xin = F.leaky_relu(x.double() * sc.double() + sh.double(), 0.2) if x_mode == 'bnact' else x.double()
    ref = dy.t() @ xin
    dw = torch.full((N, K), -0.25, device='cuda')
    t = [v.cuda() for v in (gq, yq, cA, cB, cC, x, sc, sh)]
    a = vt_dy(*t[:5]) if dy_mode == 'dy' else vt_plain(t[0])
    b = vt_bnact(t[5], t[6], t[7], L.ACT_LEAKY02) if x_mode == 'bnact' else vt_plain(t[5])
    n0 = L.lib.b200sp_launch_count()
    L.call('b200sp_pw_wgrad', C.byref(a), C.byref(b), dw.data_ptr(), None, M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert L.lib.b200sp_launch_count() - n0 == 1
    assert rel(dw + 0.25, ref) < 2e-5


@pytest.mark.parametrize('B,H,W,Cc,s', [(2, 13, 9, 32, 1), (3, 14, 14, 144, 2), (2, 7, 7, 1280, 1), (1, 33, 20, 96, 2), (2, 56, 56, 24, 1),
                                        # small maps with C % 64 == 0: the slab kernels (whole images in shared memory)
                                        (3, 14, 14, 576, 1), (3, 14, 14, 576, 2), (11, 7, 7, 960, 1), (2, 14, 14, 384, 1), (9, 7, 7, 64, 2),
                                        (48, 7, 7, 320, 1), (5, 13, 11, 128, 2), (1, 2, 3, 64, 1)])
def test_dw_fwd(B, H, W, Cc, s):
    g = _g(H * W + Cc + s)
    x = torch.randn(B, H, W, Cc, generator=g)
    w = torch.randn(Cc, 1, 3, 3, generator=g)
    sc, sh = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.5
    xin = torch.clamp(x.double() * sc.double() + sh.double(), 0, 6).permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.double(), None, s, 1, 1, Cc).permute(0, 2, 3, 1)
    Ho, Wo = ref.shape[1], ref.shape[2]
    y = torch.empty(B, Ho, Wo, Cc, device='cuda')
    w9 = w.view(Cc, 9).t().contiguous().cuda()
    bn = BnF(Cc)
    xc, scc, shc = x.cuda(), sc.cuda(), sh.cuda()
    L.call('b200sp_dw_fwd', C.byref(vt_bnact(xc, scc, shc, L.ACT_RELU6)), w9.data_ptr(), y.data_ptr(), bn.ref(),
           B, H, W, Cc, s, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(y, ref) < 1e-5
    assert rel(bn.mean, ref.mean((0, 1, 2))) < 1e-5
    assert rel(bn.rstd, 1 / torch.sqrt(ref.var((0, 1, 2), unbiased=False) + 1e-5)) < 1e-5


@pytest.mark.parametrize('B,H,W,Cc,s', [(2, 13, 9, 32, 1), (3, 14, 14, 144, 2), (2, 7, 7, 320, 1), (1, 33, 20, 96, 2), (2, 28, 28, 192, 1),
                                        (3, 14, 14, 576, 1), (3, 14, 14, 576, 2), (11, 7, 7, 960, 1), (2, 14, 14, 384, 1), (9, 7, 7, 64, 2),
                                        (5, 13, 11, 128, 2), (1, 2, 3, 64, 1)])
@pytest.mark.parametrize('with_skip', [False, True])
def test_dw_bwd_fused(B, H, W, Cc, s, with_skip):
    g = _g(H + W + Cc + s)
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    yin = torch.randn(B, H, W, Cc, generator=g) * 2              # raw output of the producer of the conv input
    sc, sh = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.5
    mean, rstd = torch.randn(Cc, generator=g) * 0.2, torch.rand(Cc, generator=g) + 0.5
    w = torch.randn(Cc, 1, 3, 3, generator=g)
    gq, yq = torch.randn(B, Ho, Wo, Cc, generator=g), torch.randn(B, Ho, Wo, Cc, generator=g)
    cA, cB, cC = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1, torch.randn(Cc, generator=g) * 0.1
    skip = torch.randn(B, H, W, Cc, generator=g) if with_skip else None
    # float64 reference through autograd
    z = (yin.double() * sc.double() + sh.double()).requires_grad_(True)
    a = torch.clamp(z, 0, 6)
    wd = w.double().requires_grad_(True)
    out = F.conv2d(a.permute(0, 3, 1, 2), wd, None, s, 1, 1, Cc).permute(0, 2, 3, 1)
    dy = cA.double() * gq.double() + cB.double() * yq.double() + cC.double()
    a.retain_grad()
    out.backward(dy)
    da = a.grad + (skip.double() if with_skip else 0)
    mask = ((z > 0) & (z < 6)).double()
    gref = da * mask
    xhat = (yin.double() - mean.double()) * rstd.double()
    d = {k: v.cuda() for k, v in dict(yin=yin, sc=sc, sh=sh, mean=mean, rstd=rstd, gq=gq, yq=yq, cA=cA, cB=cB, cC=cC).items()}
    w9 = w.view(Cc, 9).t().contiguous().cuda()
    gin = torch.empty(B, H, W, Cc, device='cuda')
    dw9 = torch.zeros(9, Cc, device='cuda')
    bn = BnB(d['yin'], d['sc'], d['sh'], d['mean'], d['rstd'], L.ACT_RELU6)
    skd = skip.cuda() if with_skip else None
    L.call('b200sp_dw_bwd', C.byref(vt_dy(d['gq'], d['yq'], d['cA'], d['cB'], d['cC'])),
           C.byref(vt_bnact(d['yin'], d['sc'], d['sh'], L.ACT_RELU6)), w9.data_ptr(), skd.data_ptr() if with_skip else None,
           gin.data_ptr(), dw9.data_ptr(), bn.ref(), B, H, W, Cc, s, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(gin, gref) < 1e-5
    assert rel(dw9.t().reshape(Cc, 1, 3, 3), wd.grad) < 2e-5
    assert rel(bn.dbeta, gref.sum((0, 1, 2))) < 1e-5
    assert rel(bn.dgamma, (gref * xhat).sum((0, 1, 2))) < 1e-4


def test_dw_bwd_plain_input_no_bn():
    g = _g(77)
    B, H, W, Cc = 2, 7, 7, 1280
    x = torch.randn(B, H, W, Cc, generator=g)
    w = torch.randn(Cc, 1, 3, 3, generator=g)
    dy = torch.randn(B, H, W, Cc, generator=g)
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    F.conv2d(xd.permute(0, 3, 1, 2), wd, None, 1, 1, 1, Cc).permute(0, 2, 3, 1).backward(dy.double())
    gin, dw9 = torch.empty(B, H, W, Cc, device='cuda'), torch.zeros(9, Cc, device='cuda')
    dyc, xc, w9 = dy.cuda(), x.cuda(), w.view(Cc, 9).t().contiguous().cuda()
    L.call('b200sp_dw_bwd', C.byref(vt_plain(dyc)), C.byref(vt_plain(xc)), w9.data_ptr(),
           None, gin.data_ptr(), dw9.data_ptr(), None, B, H, W, Cc, 1, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(gin, xd.grad) < 1e-5 and rel(dw9.t().reshape(Cc, 1, 3, 3), wd.grad) < 2e-5


@pytest.mark.parametrize('B,H,W', [(2, 224, 224), (3, 30, 18)])
def test_stem_fwd_and_wgrad(B, H, W):
    g = _g(H)
    x, w = torch.rand(B, 3, H, W, generator=g), torch.randn(32, 3, 3, 3, generator=g) * 0.3
    xd, wd = x.double(), w.double().requires_grad_(True)
    ref = F.conv2d(xd, wd, None, 2, 1)
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.empty(B, Ho, Wo, 32, device='cuda')
    bn = BnF(32)
    xc, wc = x.cuda(), w.cuda()
    L.call('b200sp_stem_fwd', xc.data_ptr(), wc.data_ptr(), y.data_ptr(), bn.ref(), B, H, W, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(y.permute(0, 3, 1, 2), ref) < 1e-5
    assert rel(bn.mean, ref.mean((0, 2, 3))) < 1e-5
    dy = torch.randn(B, Ho, Wo, 32, generator=g)
    ref.backward(dy.permute(0, 3, 1, 2).double())
    dw = torch.zeros(32, 27, device='cuda')
    dyc = dy.cuda()
    L.call('b200sp_stem_wgrad', xc.data_ptr(), C.byref(vt_plain(dyc)), dw.data_ptr(), B, H, W, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(dw.view(32, 3, 3, 3), wd.grad) < 2e-5


def test_bn_apply_residual_and_eval_affine():
    g = _g(5)
    M, Cc = 1234, 24
    y, r = torch.randn(M, Cc, generator=g), torch.randn(M, Cc, generator=g)
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g)
    rm, rv = torch.randn(Cc, generator=g), torch.rand(Cc, generator=g) + 0.5
    sc, sh = torch.empty(Cc, device='cuda'), torch.empty(Cc, device='cuda')
    gc, bc, rmc, rvc, yc, rc = gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda(), y.cuda(), r.cuda()
    L.call('b200sp_bn_eval_affine', gc.data_ptr(), bc.data_ptr(), rmc.data_ptr(), rvc.data_ptr(),
           1e-5, sc.data_ptr(), sh.data_ptr(), Cc, sp())
    out = torch.empty(M, Cc, device='cuda')
    L.call('b200sp_bn_apply', yc.data_ptr(), sc.data_ptr(), sh.data_ptr(), rc.data_ptr(), L.ACT_NONE, out.data_ptr(),
           M, Cc, L.F32, sp())
    torch.cuda.synchronize()
    ref = F.batch_norm(y.double(), rm.double(), rv.double(), gamma.double(), beta.double(), False, 0.1, 1e-5) + r.double()
    assert rel(out, ref) < 1e-6


def test_reorg_cat_matches_reference_view_transpose_chain():
    g = _g(6)
    B, h, w, Cr, C1 = 2, 7, 7, 64, 1024
    xr, x1 = torch.randn(B, 2 * h, 2 * w, Cr, generator=g), torch.randn(B, h, w, C1, generator=g)
    out = torch.empty(B, h, w, 4 * Cr + C1, device='cuda')
    xrc, x1c = xr.cuda(), x1.cuda()
    L.call('b200sp_reorg_cat_fwd', C.byref(vt_plain(xrc)), C.byref(vt_plain(x1c)), out.data_ptr(), B, h, w, Cr, C1, L.F32, sp())
    torch.cuda.synchronize()
    x2 = xr.permute(0, 3, 1, 2).contiguous()          # the reference's NCHW chain, park2019.py:74-80
    Bc, Cc, H, W = x2.shape
    s = 2
    x2 = x2.view(Bc, Cc, H // s, s, W // s, s).transpose(3, 4).contiguous()
    x2 = x2.view(Bc, Cc, H // s * W // s, s * s).transpose(2, 3).contiguous()
    x2 = x2.view(Bc, Cc, s * s, H // s, W // s).transpose(1, 2).contiguous()
    x2 = x2.view(Bc, s * s * Cc, H // s, W // s)
    ref = torch.cat((x2, x1.permute(0, 3, 1, 2)), 1)
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), ref)


def test_head_fwd_loss_bwd():
    g = _g(8)
    B, Cc, N = 5, 1024, 22
    HWC = 49 * Cc
    y = torch.randn(B, 7, 7, Cc, generator=g)
    sc, sh = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.5
    w = torch.randn(N, Cc, 7, 7, generator=g) / 200
    bias, tgt = torch.randn(N, generator=g), torch.rand(B, 2, 11, generator=g)
    z = (y.double() * sc.double() + sh.double()).requires_grad_(True)
    a = torch.relu(z)
    wd, bd = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    logits = F.conv2d(a.permute(0, 3, 1, 2), wd, bd).view(B, N)
    loss = sum(F.mse_loss(logits[:, 2 * i], tgt[:, 0, i].double()) + F.mse_loss(logits[:, 2 * i + 1], tgt[:, 1, i].double()) for i in range(11))
    loss.backward()
    yc, scc, shc = y.cuda(), sc.cuda(), sh.cuda()
    wn = w.permute(0, 2, 3, 1).contiguous().cuda()
    lg, dl, l3 = torch.empty(B, N, device='cuda'), torch.empty(B, N, device='cuda'), torch.empty(3, device='cuda')
    vt = vt_bnact(yc, scc, shc, L.ACT_RELU)
    biasc, tgtc = bias.cuda(), tgt.cuda()
    L.call('b200sp_head_bias', biasc.data_ptr(), lg.data_ptr(), B, N, sp())
    L.call('b200sp_head_fwd', C.byref(vt), wn.data_ptr(), lg.data_ptr(), B, HWC, Cc, N, L.F32, sp())
    L.call('b200sp_krn_loss', lg.data_ptr(), tgtc.data_ptr(), l3.data_ptr(), dl.data_ptr(), None, None, B, N, sp())
    torch.cuda.synchronize()
    assert rel(lg, logits) < 1e-5 and abs(float(l3[0]) - float(loss)) < 1e-5 * float(loss)
    gq, dwn, db = torch.empty(B, 7, 7, Cc, device='cuda'), torch.zeros(N, 7, 7, Cc, device='cuda'), torch.zeros(N, device='cuda')
    mean, rstd = torch.zeros(Cc, device='cuda'), torch.ones(Cc, device='cuda')
    bn = BnB(yc, scc, shc, mean, rstd, L.ACT_RELU)
    L.call('b200sp_head_bwd', dl.data_ptr(), C.byref(vt), wn.data_ptr(), gq.data_ptr(), dwn.data_ptr(), db.data_ptr(), bn.ref(),
           B, HWC, Cc, N, L.F32, sp())
    torch.cuda.synchronize()
    assert rel(gq, z.grad) < 1e-5
    assert rel(dwn.permute(0, 3, 1, 2), wd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5
    assert rel(bn.dbeta, z.grad.sum((0, 1, 2))) < 1e-5


@pytest.mark.parametrize('clip_mode', [0, 1, 2])
def test_adamw_matches_oracle(clip_mode):
    from speedplusbaseline_b200.params import ParamStore
    from speedplusbaseline_b200.optim import FusedAdamW
    g = _g(11 + clip_mode)
    st = ParamStore([('a.weight', 'plain', (37, 5)), ('b.weight', 'plain', (1001,))], [('bn', 8)], torch.device('cuda'))
    p0 = torch.randn(st.n, generator=g)
    st.params.copy_(p0)
    opt = FusedAdamW(st, [torch.nn.Parameter(st.params)], lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=clip_mode)
    ref_p = [p0.clone().double()]
    ost = ooptim.AdamWState(ref_p)
    for it in range(3):
        gr = torch.randn(st.n, generator=g) * (3.0 if it == 1 else 0.01)
        st.grads.copy_(gr)
        opt.step()
        gl = [gr.clone().double()]
        if clip_mode == 1:
            ooptim.clip_grad_norm(gl, 1.0)
        elif clip_mode == 2:
            ooptim.clip_grad_value(gl, 1.0)
        ooptim.adamw_step(ref_p, gl, ost, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01)
        torch.cuda.synchronize()
        assert rel(st.params, ref_p[0]) < 2e-6, it
        if clip_mode == 1:
            assert abs(opt.last_grad_norm() - float(gr.double().norm())) < 1e-4 * float(gr.norm())
