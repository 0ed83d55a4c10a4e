"""torch.library registration of the libb200sp kernels: `torch.ops.b200sp.*` (north_star: "registered as a torch extension";
SURVEY.md 8b lower side).  The operators take and return torch tensors (NHWC activations, fp32), run on the current CUDA
stream, never synchronise and allocate only through torch's caching allocator.  They are thin adapters over the C-ABI of
include/b200sp.h -- the same entry points the engines call through ctypes -- registered with the dispatcher for the CUDA key
only, so calling one with CPU tensors fails loudly (there is no CPU fallback anywhere in this package).

Why the schemas are declared here with torch.library instead of a TORCH_LIBRARY block in C++: the product boundary is the plain
C-ABI shared object (no torch types in any signature); compiling a second, torch-linked translation unit
would add a libtorch build dependency for the same dispatcher entries this module creates in ~100 lines.

    import speedplusbaseline_b200.torch_ops          # registers the namespace
    y, mean, rstd, scale, shift = torch.ops.b200sp.conv1x1_fwd(x, w, None, None, 0, gamma, beta, 1e-5)

Operators (each cites the reference call site it replaces):
  conv1x1_fwd / conv1x1_dgrad / conv1x1_wgrad      nn.Conv2d(k=1) fwd / input-grad / weight-grad  (park2019.py:45-54, mobilenetv2.py)
  conv_dw3x3_fwd                                   nn.Conv2d(groups=C, k=3, p=1)                  (park2019.py:46)
  bn_apply                                         BatchNorm affine (+residual) (+activation)      (mobilenetv2.py:61-62)
  reorg_cat                                        RouterV2 space-to-depth + concat               (park2019.py:70-80)
  krn_loss                                         sum of per-keypoint MSE + its gradient         (park2019.py:142-160)
  grad_sqnorm / adamw_fused                        clip_grad_norm_ + torch.optim.AdamW.step        (trainer.py:97, build.py:72-74)
"""
import ctypes as C

import torch

from . import _lib as L

_lib = torch.library.Library('b200sp', 'DEF')
_impl = torch.library.Library('b200sp', 'IMPL', 'CUDA')


def _vt(x, scale, shift, act):
    if scale is None:
        return L.VTensor(x.data_ptr(), None, None, None, None, L.VT_PLAIN, 0)
    return L.VTensor(x.data_ptr(), None, scale.data_ptr(), shift.data_ptr(), None, L.VT_BNACT, act)


def _chk(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'b200sp ops take contiguous fp32 CUDA tensors'


class _BnWork:
    """workspace of the fused BatchNorm-statistics epilogue (sum / sumsq / ticket + outputs)"""

    def __init__(self, N, gamma, beta, eps, dev):
        f = dict(device=dev, dtype=torch.float32)
        self.acc = torch.zeros(2, N, device=dev, dtype=torch.float64)
        self.ticket = torch.zeros(4, device=dev, dtype=torch.int32)
        self.out = torch.empty(4, N, **f)                      # scale shift mean rstd
        o = self.out
        self.s = L.BnFwd(self.acc[0].data_ptr(), self.acc[1].data_ptr(), self.ticket.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                         None, None, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), 0.1, float(eps))


# ---- conv1x1 -----------------------------------------------------------------------------------------------------------
_lib.define('conv1x1_fwd(Tensor x, Tensor w, Tensor? in_scale, Tensor? in_shift, int in_act, Tensor? gamma, Tensor? beta, float eps) '
            '-> (Tensor, Tensor, Tensor, Tensor, Tensor)')


def conv1x1_fwd(x, w, in_scale, in_shift, in_act, gamma, beta, eps):
    """y[M,N] = act(in_scale*x + in_shift)[M,K] @ w[N,K]^T (x raw: the producer's BatchNorm+activation is applied on load).
    With gamma/beta: also the batch statistics of y fused in the epilogue -> (y, mean, rstd, scale, shift); else those are empty."""
    _chk(x, w, in_scale, in_shift, gamma, beta)
    K = x.shape[-1]
    M, N = x.numel() // K, w.shape[0]
    y = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=torch.float32)
    vt = _vt(x, in_scale, in_shift, in_act)
    bn = _BnWork(N, gamma, beta, eps, x.device) if gamma is not None else None
    L.call('b200sp_pw_fwd', C.byref(vt), w.data_ptr(), None, 0, y.data_ptr(), C.byref(bn.s) if bn else None, M, N, K, L.F32, L.stream_ptr())
    if bn is None:
        e = torch.empty(0, device=x.device)
        return y, e, e, e, e
    o = bn.out
    return y, o[2], o[3], o[0], o[1]


_impl.impl('conv1x1_fwd', conv1x1_fwd)

_lib.define('conv1x1_dgrad(Tensor dy, Tensor w) -> Tensor')


def conv1x1_dgrad(dy, w):
    """dx[M,K] = dy[M,N] @ w[N,K]"""
    _chk(dy, w)
    N, K = w.shape[0], w.shape[1]
    M = dy.numel() // N
    dx = torch.empty(dy.shape[:-1] + (K,), device=dy.device, dtype=torch.float32)
    vt = _vt(dy, None, None, 0)
    L.call('b200sp_pw_dgrad', C.byref(vt), w.data_ptr(), None, 1.0, dx.data_ptr(), None, M, N, K, L.F32, L.stream_ptr())
    return dx


_impl.impl('conv1x1_dgrad', conv1x1_dgrad)

_lib.define('conv1x1_wgrad(Tensor dy, Tensor x, Tensor? in_scale, Tensor? in_shift, int in_act) -> Tensor')


def conv1x1_wgrad(dy, x, in_scale, in_shift, in_act):
    """dw[N,K] = dy[M,N]^T @ act(in_scale*x + in_shift)[M,K]"""
    _chk(dy, x, in_scale, in_shift)
    N, K = dy.shape[-1], x.shape[-1]
    M = x.numel() // K
    dw = torch.zeros(N, K, device=x.device, dtype=torch.float32)
    a, b = _vt(dy, None, None, 0), _vt(x, in_scale, in_shift, in_act)
    L.call('b200sp_pw_wgrad', C.byref(a), C.byref(b), dw.data_ptr(), None, M, N, K, L.F32, L.stream_ptr())
    return dw


_impl.impl('conv1x1_wgrad', conv1x1_wgrad)

# ---- depthwise 3x3 -----------------------------------------------------------------------------------------------------
_lib.define('conv_dw3x3_fwd(Tensor x, Tensor w, int stride, Tensor? in_scale, Tensor? in_shift, int in_act) -> Tensor')


def conv_dw3x3_fwd(x, w, stride, in_scale, in_shift, in_act):
    """x [B,H,W,C] NHWC, w [C,1,3,3] (reference layout) -> y [B,Ho,Wo,C], padding 1, no bias"""
    _chk(x, w, in_scale, in_shift)
    B, H, W, Cc = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    w9 = w.reshape(Cc, 9).t().contiguous()
    y = torch.empty(B, Ho, Wo, Cc, device=x.device, dtype=torch.float32)
    vt = _vt(x, in_scale, in_shift, in_act)
    L.call('b200sp_dw_fwd', C.byref(vt), w9.data_ptr(), y.data_ptr(), None, B, H, W, Cc, stride, L.F32, L.stream_ptr())
    return y


_impl.impl('conv_dw3x3_fwd', conv_dw3x3_fwd)

# ---- BatchNorm apply / RouterV2 / loss -------------------------------------------------------------------------------------
_lib.define('bn_apply(Tensor y, Tensor scale, Tensor shift, Tensor? residual, int act) -> Tensor')


def bn_apply(y, scale, shift, residual, act):
    """out = act(y*scale + shift) + residual  (the inverted-residual block output: no activation after the add)"""
    _chk(y, scale, shift, residual)
    Cc = y.shape[-1]
    out = torch.empty_like(y)
    L.call('b200sp_bn_apply', y.data_ptr(), scale.data_ptr(), shift.data_ptr(), L.ptr(residual), act, out.data_ptr(), y.numel() // Cc, Cc,
           L.F32, L.stream_ptr())
    return out


_impl.impl('bn_apply', bn_apply)

_lib.define('reorg_cat(Tensor xr, Tensor x1) -> Tensor')


def reorg_cat(xr, x1):
    """RouterV2 tail: space-to-depth(2) of xr [B,2h,2w,Cr] concatenated in front of x1 [B,h,w,C1] -> [B,h,w,4Cr+C1]"""
    _chk(xr, x1)
    B, h, w, C1 = x1.shape
    Cr = xr.shape[-1]
    out = torch.empty(B, h, w, 4 * Cr + C1, device=x1.device, dtype=torch.float32)
    a, b = _vt(xr, None, None, 0), _vt(x1, None, None, 0)
    L.call('b200sp_reorg_cat_fwd', C.byref(a), C.byref(b), out.data_ptr(), B, h, w, Cr, C1, L.F32, L.stream_ptr())
    return out


_impl.impl('reorg_cat', reorg_cat)

_lib.define('krn_loss(Tensor logits, Tensor target) -> (Tensor, Tensor)')


def krn_loss(logits, target):
    """logits [B,2K] interleaved (x0,y0,x1,..), target [B,2,K] -> (loss3 = (loss, loss_x, loss_y), dlogits)"""
    _chk(logits, target)
    B, N = logits.shape
    loss3 = torch.zeros(3, device=logits.device)
    dl = torch.empty_like(logits)
    L.call('b200sp_krn_loss', logits.data_ptr(), target.data_ptr(), loss3.data_ptr(), dl.data_ptr(), None, None, B, N, L.stream_ptr())
    return loss3, dl


_impl.impl('krn_loss', krn_loss)

# ---- optimizer ---------------------------------------------------------------------------------------------------------------
_lib.define('adamw_fused(Tensor(a!) p, Tensor g, Tensor(b!) m, Tensor(c!) v, float lr, float beta1, float beta2, float eps, '
            'float weight_decay, int step, float max_norm) -> Tensor')


def adamw_fused(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, max_norm):
    """One global-norm clip (max_norm <= 0: none) + AdamW update over FLAT buffers, in place; `step` = steps already taken.
    Returns the total gradient norm (1-element tensor)."""
    _chk(p, g, m, v)
    hp = L.AdamWHp()
    hp.lr, hp.beta1, hp.beta2, hp.eps, hp.weight_decay = lr, beta1, beta2, eps, weight_decay
    hp.max_norm, hp.clip_value, hp.grad_scale, hp.step = max(max_norm, 0.0), 1.0, 1.0, int(step)
    hp.clip_mode = 1 if max_norm > 0 else 0
    host = torch.frombuffer(bytearray(bytes(hp)), dtype=torch.uint8)
    dev = host.to(p.device)
    sp = L.stream_ptr()
    if hp.clip_mode == 1:
        L.call('b200sp_grad_sqnorm', g.data_ptr(), g.numel(), dev.data_ptr(), sp)
    L.call('b200sp_adamw_step', p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), None, p.numel(), dev.data_ptr(), sp)
    off = L.AdamWHp.last_norm.offset
    return dev[off:off + 4].view(torch.float32).clone()


_impl.impl('adamw_fused', adamw_fused)

OPS = ('conv1x1_fwd', 'conv1x1_dgrad', 'conv1x1_wgrad', 'conv_dw3x3_fwd', 'bn_apply', 'reorg_cat', 'krn_loss', 'adamw_fused')
