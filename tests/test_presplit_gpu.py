"""Presplit route of the 1x1-convolution GEMMs (tcgemm2.cu PRE mode + opsplit.cu): layers whose activation operand has at most
4096 rows (the 7x7 maps of the KRN at batch 48: park2019.py:100-118) split their operands once per call into tf32 hi/lo planes and
run a converter-free TMA -> tcgen05 kernel.  Same float64 checks as the general kernel (test_kernels_gpu.py), on shapes that take
the route, plus the proof that the route WAS taken (two launches per call: pre-pass + GEMM)."""
import ctypes as C

import pytest
import torch

from speedplusbaseline_b200 import _lib as L
from kutil import BnF, rel, sp, vt_bnact, vt_plain
import test_kernels_gpu as TK

pytestmark = pytest.mark.gpu

SHAPES = [(2352, 1024, 1280), (2352, 160, 960), (2352, 960, 160), (2352, 320, 960), (2352, 1280, 320), (2352, 160, 576),
          (1000, 200, 136), (4000, 72, 132), (260, 64, 128), (588, 1024, 1024)]


@pytest.fixture(autouse=True)
def _workspace():
    L.ensure_workspace('cuda:0')


def _launches():
    return L.lib.b200sp_launch_count()


def _eligible(kind, M, N, K):
    """tcgemm2_presplit's rule: activation rows in [256, 4096], reduction >= 128, both output extents >= 64"""
    P, Q, R = {'fwd': (M, N, K), 'dgrad': (M, K, N), 'wgrad': (N, K, M)}[kind]
    return 256 <= M <= 4096 and R >= 128 and P >= 64 and Q >= 64


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('act', [L.ACT_NONE, L.ACT_RELU6])
def test_fwd(M, N, K, act):
    n0 = _launches()
    TK.test_pw_fwd_with_bn_epilogue(M, N, K, act)
    assert _launches() - n0 == (2 if _eligible('fwd', M, N, K) else 1), 'presplit route not taken'


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('act', [L.ACT_NONE, L.ACT_RELU6, L.ACT_LEAKY02])
def test_dgrad(M, N, K, act):
    n0 = _launches()
    TK.test_pw_dgrad_fused_bn_backward(M, N, K, act)
    assert _launches() - n0 == (2 if _eligible('dgrad', M, N, K) else 1), 'presplit route not taken'


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_wgrad(M, N, K):
    n0 = _launches()
    TK.test_pw_wgrad(M, N, K)
    assert _launches() - n0 >= 2


def test_routes_agree_and_single_pass_tf32():
    """the same call with and without the workspace: both are 3xTF32 with the same split, so they agree to accumulation order;
    the --use_fp16 mode (single-pass TF32, one plane) stays within TF32's 2^-11 operand rounding"""
    g = torch.Generator().manual_seed(5)
    M, N, K = 2352, 320, 960
    x, w = torch.randn(M, K, generator=g).cuda(), (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    sc, sh = (torch.rand(K, generator=g) + 0.5).cuda(), (torch.randn(K, generator=g) * 0.3).cuda()
    ref = torch.clamp(x.double() * sc.double() + sh.double(), 0, 6) @ w.double().t()
    outs = {}
    for name, ws, dt in (('pre', True, L.F32), ('gen', False, L.F32), ('pre_tf32', True, L.F32_TF32X1), ('gen_tf32', False, L.F32_TF32X1)):
        if ws:
            L.ensure_workspace('cuda:0')
        else:
            L.call('b200sp_set_workspace', None, 0)
        y = torch.empty(M, N, device='cuda')
        bn = BnF(N, gamma=torch.ones(N, device='cuda'), beta=torch.zeros(N, device='cuda'))
        n0 = _launches()
        L.call('b200sp_pw_fwd', C.byref(vt_bnact(x, sc, sh, L.ACT_RELU6)), w.data_ptr(), None, 0, y.data_ptr(), bn.ref(), M, N, K, dt, sp())
        torch.cuda.synchronize()
        assert _launches() - n0 == (2 if ws else 1)
        outs[name] = y
    L.ensure_workspace('cuda:0')
    assert rel(outs['pre'], ref) < 2e-5 and rel(outs['gen'], ref) < 2e-5
    assert rel(outs['pre'], outs['gen'].double()) < 2e-6
    assert rel(outs['pre_tf32'], ref) < 3e-3 and rel(outs['pre_tf32'], outs['gen_tf32'].double()) < 1e-5


def test_no_workspace_is_the_general_kernel_not_an_error():
    L.call('b200sp_set_workspace', None, 0)
    g = torch.Generator().manual_seed(6)
    M, N, K = 2352, 160, 576
    x, w = torch.randn(M, K, generator=g).cuda(), (torch.randn(N, K, generator=g) / 24).cuda()
    y = torch.empty(M, N, device='cuda')
    n0 = _launches()
    L.call('b200sp_pw_fwd', C.byref(vt_plain(x)), w.data_ptr(), None, 0, y.data_ptr(), None, M, N, K, L.F32, sp())
    torch.cuda.synchronize()
    assert _launches() - n0 == 1
    assert rel(y, x.double() @ w.double().t()) < 2e-5
    with pytest.raises(L.B200SPError):
        L.call('b200sp_set_workspace', 128, 1 << 20)        # not 256-byte aligned
    L.ensure_workspace('cuda:0')
