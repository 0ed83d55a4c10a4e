"""KRN / RevGrad execution engine: the layer plan, HBM buffers and kernel sequencing.

Mirrors the graph of /root/reference/src/nets/park2019.py:101-165 (MobileNetV2 features[:-1] +
ConvDw extras + RouterV2 + 7x7 head) and src/nets/revgrad.py:58-96 (domain classifier), but every
op is a libb200sp launch (include/b200sp.h).  Design:

  * activations NHWC; each conv stores only its RAW output Y once.  The BatchNorm+activation the
    reference materialises is folded into the NEXT kernel's operand load (b200sp_vtensor), the
    BatchNorm statistics into the producing conv's epilogue;
  * backward keeps g = dL/d(act out) * act'(z) per layer; the BatchNorm backward
    (dy = cA*g + cB*y + cC) is again folded into the consumer's load, its reductions into the
    producer's epilogue;
  * all parameters / gradients live in one flat buffer (params.ParamStore) so clip + AdamW are two launches.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from .params import ParamStore

# torchvision mobilenetv2.py:105-114 inverted-residual table (t, c, n, s)
_MBV2 = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
BN_EPS, BN_MOM = 1e-5, 0.1


def _blocks():
    out, cin, idx = [], 32, 1
    for t, c, n, s in _MBV2:
        for i in range(n):
            out.append(dict(idx=idx, cin=cin, cout=c, stride=s if i == 0 else 1, t=t, res=False))
            cin, idx = c, idx + 1
    for b in out:
        b['res'] = b['stride'] == 1 and b['cin'] == b['cout']
    return out


def krn_layout(num_keypoints=11, prefix='', dann=False):
    """-> (weights [(key, kind, ref_shape)], bns [(prefix, C)], state_dict key order)."""
    W, BN, order = [], [], []

    def conv(key, kind, shp):
        W.append((prefix + key, kind, shp))
        order.append(prefix + key)

    def bn(p, c):
        BN.append((prefix + p, c))
        for f in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked'):
            order.append(prefix + p + '.' + f)

    conv('base.0.0.weight', 'plain', (32, 3, 3, 3)); bn('base.0.1', 32)
    for b in _blocks():
        p = 'base.%d.conv' % b['idx']
        hid, j = b['cin'] * b['t'], 0
        if b['t'] != 1:
            conv(p + '.0.0.weight', 'plain', (hid, b['cin'], 1, 1)); bn(p + '.0.1', hid)
            j = 1
        conv('%s.%d.0.weight' % (p, j), 'dw', (hid, 1, 3, 3)); bn('%s.%d.1' % (p, j), hid)
        conv('%s.%d.weight' % (p, j + 1), 'plain', (b['cout'], hid, 1, 1)); bn('%s.%d' % (p, j + 2), b['cout'])
    for i, (inp, oup) in ((0, (320, 1024)), (1, (1024, 1024)), (2, (96, 64)), (3, (1280, 1024))):
        p = 'extras.%d.conv' % i
        if i == 2:
            conv(p + '.0.weight', 'plain', (64, 96, 1, 1)); bn(p + '.1', 64)
        else:
            conv(p + '.0.weight', 'dw', (inp, 1, 3, 3)); bn(p + '.1', inp)
            conv(p + '.3.weight', 'plain', (oup, inp, 1, 1)); bn(p + '.4', oup)
    conv('head.0.weight', 'ohwi', (2 * num_keypoints, 1024, 7, 7))
    conv('head.0.bias', 'plain', (2 * num_keypoints,))
    if dann:
        for k, shp in (('domain_classifier.0.weight', (1280, 320, 1, 1)), ('domain_classifier.0.bias', (1280,)),
                       ('domain_classifier.3.weight', (1, 1280, 1, 1)), ('domain_classifier.3.bias', (1,))):
            W.append((k, 'plain', shp))
            order.append(k)
    return W, BN, order


class _Ctx:
    """Per-forward-pass HBM state (activations, BN statistics, gradients wrt activations)."""

    def __init__(self, eng, B):
        dev, st = eng.device, eng.store
        self.B = B
        f32 = dict(dtype=torch.float32, device=dev)
        self.stat = torch.zeros(7, st.totC, **f32)            # scale shift mean rstd cA cB cC
        self.sums = torch.zeros(4, st.totC, dtype=torch.float64, device=dev)   # sum sumsq s1 s2
        self.tick = torch.zeros(2 * len(st.bns) + 8, dtype=torch.int32, device=dev)
        self.Y, self.G, self.O, self.dO = {}, {}, {}, {}
        self.logits = torch.zeros(B, eng.N, **f32)
        self.dlogits = torch.zeros(B, eng.N, **f32)
        self.loss3 = torch.zeros(3, **f32)
        self.keep = []       # ctypes structs referenced by in-flight launches

    def sp(self, row, bn_i, st):
        return self.stat.data_ptr() + 4 * (row * st.totC + st.bn_off[bn_i])

    def dp(self, row, bn_i, st):
        return self.sums.data_ptr() + 8 * (row * st.totC + st.bn_off[bn_i])


import os as _os
_POISON = _os.environ.get('B200SP_DEBUG_POISON') == '1'


class KRNEngine:
    def __init__(self, num_keypoints=11, prefix='', dann=False, device=None, dtype=L.F32, tf32_gemm=False):
        """tf32_gemm: the `--use_fp16` mode -- fp32 storage everywhere, 1x1-convolution GEMMs in SINGLE-pass TF32 (fp16's 10-bit
        operand mantissa with fp32 range and accumulation; include/b200sp.h B200SP_F32_TF32X1)."""
        L.require_cuda()
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.prefix, self.dann, self.dtype = prefix, dann, dtype
        self.gemm_dtype = L.F32_TF32X1 if (tf32_gemm and dtype == L.F32) else dtype
        self.nk, self.N = num_keypoints, 2 * num_keypoints
        W, BN, self.key_order = krn_layout(num_keypoints, prefix, dann)
        self.store = ParamStore(W, BN, self.device)
        self.blocks = _blocks()
        self._ctxs = {}
        L.ensure_workspace(self.device)       # presplit route of the 7x7-layer GEMMs (tcgemm2.cu PRE mode)
        self.tdtype = torch.float32 if dtype == L.F32 else torch.bfloat16
        import os
        self._async_wgrad = os.environ.get('B200SP_ASYNC_WGRAD', '1') != '0'
        self._side = torch.cuda.Stream(device=self.device) if self._async_wgrad else None
        self._side_used = False
        if dtype == L.BF16:
            self.store.enable_lowp()

    # ------------------------------------------------------------------ helpers
    def ctx(self, B, slot=0):
        k = (B, slot)
        if k not in self._ctxs:
            self._ctxs[k] = _Ctx(self, B)
        return self._ctxs[k]

    def _buf(self, d, name, shape):
        t = d.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=self.tdtype, device=self.device)
            if _POISON:          # B200SP_DEBUG_POISON=1: a kernel that reads a buffer before its producer wrote it shows up as NaN
                t.fill_(float('nan'))
            d[name] = t
        return t

    def _bi(self, p):
        return self.store.bn_index[self.prefix + p]

    def _vt_plain(self, t):
        return L.VTensor(t.data_ptr(), None, None, None, None, L.VT_PLAIN, 0)

    def _vt_bnact(self, cx, y, bn_i, act):
        st = self.store
        return L.VTensor(y.data_ptr(), None, cx.sp(0, bn_i, st), cx.sp(1, bn_i, st), None, L.VT_BNACT, act)

    def _vt_dy(self, cx, g, y, bn_i):
        st = self.store
        return L.VTensor(g.data_ptr(), y.data_ptr(), cx.sp(4, bn_i, st), cx.sp(5, bn_i, st), cx.sp(6, bn_i, st), L.VT_DY, 0)

    def _bnfwd(self, cx, bn_i, train):
        if not train:
            return None
        st = self.store
        g, b, rm, rv = st.bn_slices(bn_i)
        s = L.BnFwd(cx.dp(0, bn_i, st), cx.dp(1, bn_i, st), cx.tick.data_ptr() + 4 * (2 * bn_i),
                    st.p_ptr(g.start), st.p_ptr(b.start),
                    st.bufs.data_ptr() + 4 * rm.start, st.bufs.data_ptr() + 4 * rv.start,
                    cx.sp(0, bn_i, st), cx.sp(1, bn_i, st), cx.sp(2, bn_i, st), cx.sp(3, bn_i, st), BN_MOM, BN_EPS)
        cx.keep.append(s)
        return C.byref(s)

    def _bnbwd(self, cx, bn_i, y, act, stats=True):
        st = self.store
        g, b, _, _ = st.bn_slices(bn_i)
        s = L.BnBwd(cx.dp(2, bn_i, st) if stats else None, cx.dp(3, bn_i, st) if stats else None,
                    cx.tick.data_ptr() + 4 * (2 * bn_i + 1), y.data_ptr(),
                    cx.sp(0, bn_i, st), cx.sp(1, bn_i, st), cx.sp(2, bn_i, st), cx.sp(3, bn_i, st),
                    cx.sp(4, bn_i, st), cx.sp(5, bn_i, st), cx.sp(6, bn_i, st),
                    st.g_ptr(g.start), st.g_ptr(b.start), act, 0)
        cx.keep.append(s)
        return C.byref(s)

    def _w(self, key):
        return self.store.w_ptr(self.prefix + key)

    def _wg(self, key):
        return self.store.wg_ptr(self.prefix + key)

    def _wgrad(self, *args):
        """b200sp_pw_wgrad on the side stream: a layer's weight gradient and its data gradient are independent
        consumers of dY, and each of these GEMMs leaves most SMs idle (few tiles, latency-bound), so they overlap.
        Fork after the kernel that produced dY (event), join at the end of backward().  Works under graph capture."""
        args = args[:-1] + (self.gemm_dtype,)
        if not self._async_wgrad:
            L.call('b200sp_pw_wgrad', *args, L.stream_ptr())
            return
        ev = torch.cuda.Event()
        ev.record()
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            L.call('b200sp_pw_wgrad', *args, L.stream_ptr())
        self._side_used = True

    def _join_side(self):
        if self._async_wgrad and self._side_used:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_used = False

    def _wq(self, key):
        """GEMM weight operand in the activation dtype (bf16 mirror for --use_fp16, the fp32 master otherwise)."""
        if self.dtype == L.BF16:
            return self.store.wl_ptr(self.prefix + key)
        return self.store.w_ptr(self.prefix + key)

    # ------------------------------------------------------------------ layer launchers
    def _pw_fwd(self, cx, xvt, wkey, bn_p, M, N, K, train, name, bias=None, out_act=L.ACT_NONE, shape=None):
        y = self._buf(cx.Y, name, shape)
        bn = self._bnfwd(cx, self._bi(bn_p), train) if bn_p else None
        L.call('b200sp_pw_fwd', C.byref(xvt), self._wq(wkey), bias, out_act, y.data_ptr(), bn, M, N, K, self.gemm_dtype, L.stream_ptr())
        return y

    def _dw_fwd(self, cx, xvt, wkey, bn_p, B, H, W, Cc, stride, train, name):
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = self._buf(cx.Y, name, (B, Ho, Wo, Cc))
        L.call('b200sp_dw_fwd', C.byref(xvt), self._w(wkey), y.data_ptr(), self._bnfwd(cx, self._bi(bn_p), train),
               B, H, W, Cc, stride, self.dtype, L.stream_ptr())
        return y

    # ------------------------------------------------------------------ forward
    def forward(self, images, target=None, train=True, slot=0):
        """images [B,3,H,W] fp32 NCHW on device.  Returns the context (logits, loss3 filled).
        train=True uses batch statistics and updates running stats (module.train() semantics)."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        B, _, H, W = images.shape
        cx = self.ctx(B, slot)
        cx.keep.clear()
        cx.images, cx.H, cx.W, cx.train = images, H, W, train
        st, sp, dt = self.store, L.stream_ptr(), self.dtype
        if train:
            L.call('b200sp_add_i64', st.nbt.data_ptr(), len(st.bns), 1, sp)
        else:
            L.call('b200sp_bn_eval_affine', st.p_ptr(st.gamma_off), st.p_ptr(st.beta_off), st.bufs.data_ptr(),
                   st.bufs.data_ptr() + 4 * st.totC, BN_EPS, cx.stat.data_ptr(), cx.stat.data_ptr() + 4 * st.totC, st.totC, sp)
        h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y0 = self._buf(cx.Y, 'stem', (B, h, w, 32))
        L.call('b200sp_stem_fwd', images.data_ptr(), self._w('base.0.0.weight'), y0.data_ptr(),
               self._bnfwd(cx, self._bi('base.0.1'), train), B, H, W, dt, sp)
        xvt = self._vt_bnact(cx, y0, self._bi('base.0.1'), L.ACT_RELU6)
        cx.geo = {}
        prevO = None
        for b in self.blocks:
            i, cin, cout, s, t = b['idx'], b['cin'], b['cout'], b['stride'], b['t']
            p = 'base.%d.conv' % i
            hid, j = cin * t, 0
            cx.geo[i] = (h, w)
            if t != 1:
                ye = self._pw_fwd(cx, xvt, p + '.0.0.weight', p + '.0.1', B * h * w, hid, cin, train, 'e%d' % i, shape=(B, h, w, hid))
                xvt = self._vt_bnact(cx, ye, self._bi(p + '.0.1'), L.ACT_RELU6)
                j = 1
            yd = self._dw_fwd(cx, xvt, '%s.%d.0.weight' % (p, j), '%s.%d.1' % (p, j), B, h, w, hid, s, train, 'd%d' % i)
            h, w = (h - 1) // s + 1, (w - 1) // s + 1
            dvt = self._vt_bnact(cx, yd, self._bi('%s.%d.1' % (p, j)), L.ACT_RELU6)
            bnp = '%s.%d' % (p, j + 2)
            yp = self._pw_fwd(cx, dvt, '%s.%d.weight' % (p, j + 1), bnp, B * h * w, cout, hid, train, 'p%d' % i, shape=(B, h, w, cout))
            o = self._buf(cx.O, i, (B, h, w, cout))
            bi = self._bi(bnp)
            L.call('b200sp_bn_apply', yp.data_ptr(), cx.sp(0, bi, st), cx.sp(1, bi, st),
                   prevO.data_ptr() if b['res'] else None, L.ACT_NONE, o.data_ptr(), B * h * w, cout, dt, sp)
            prevO = o
            xvt = self._vt_plain(o)
        cx.fh, cx.fw = h, w                                  # feature map size (7x7 at 224)
        M = B * h * w
        # extras.0 / extras.1 : ConvDw (park2019.py:45-54)
        cur_vt, cur_c = self._vt_plain(cx.O[17]), 320
        for e, oup in ((0, 1024), (1, 1024)):
            p = 'extras.%d.conv' % e
            yd = self._dw_fwd(cx, cur_vt, p + '.0.weight', p + '.1', B, h, w, cur_c, 1, train, 'xd%d' % e)
            dvt = self._vt_bnact(cx, yd, self._bi(p + '.1'), L.ACT_RELU)
            yp = self._pw_fwd(cx, dvt, p + '.3.weight', p + '.4', M, oup, cur_c, train, 'xp%d' % e, shape=(B, h, w, oup))
            cur_vt, cur_c = self._vt_bnact(cx, yp, self._bi(p + '.4'), L.ACT_RELU), oup
        # extras.2 : RouterV2 on base[13] output (park2019.py:70-80)
        h13, w13 = cx.O[13].shape[1], cx.O[13].shape[2]
        yr = self._pw_fwd(cx, self._vt_plain(cx.O[13]), 'extras.2.conv.0.weight', 'extras.2.conv.1', B * h13 * w13, 64, 96,
                          train, 'xr', shape=(B, h13, w13, 64))
        rvt = self._vt_bnact(cx, yr, self._bi('extras.2.conv.1'), L.ACT_LEAKY02)
        cat = self._buf(cx.O, 'cat', (B, h, w, 1280))
        L.call('b200sp_reorg_cat_fwd', C.byref(rvt), C.byref(cur_vt), cat.data_ptr(), B, h, w, 64, 1024, dt, sp)
        # extras.3
        p = 'extras.3.conv'
        yd = self._dw_fwd(cx, self._vt_plain(cat), p + '.0.weight', p + '.1', B, h, w, 1280, 1, train, 'xd3')
        dvt = self._vt_bnact(cx, yd, self._bi(p + '.1'), L.ACT_RELU)
        yp = self._pw_fwd(cx, dvt, p + '.3.weight', p + '.4', M, 1024, 1280, train, 'xp3', shape=(B, h, w, 1024))
        hvt = self._vt_bnact(cx, yp, self._bi(p + '.4'), L.ACT_RELU)
        cx.head_vt = hvt
        # head (park2019.py:121,139): valid 7x7 conv == FC over the flattened NHWC map
        assert h == 7 and w == 7, 'KRN head is a 7x7 valid conv: input must be 224x224'
        L.call('b200sp_head_bias', self._w('head.0.bias'), cx.logits.data_ptr(), B, self.N, sp)
        L.call('b200sp_head_fwd', C.byref(hvt), self._w('head.0.weight'), cx.logits.data_ptr(), B, h * w * 1024, 1024, self.N, dt, sp)
        cx.has_loss = target is not None
        if target is not None:
            assert target.is_cuda and target.dtype == torch.float32 and target.is_contiguous()
            L.call('b200sp_krn_loss', cx.logits.data_ptr(), target.data_ptr(), cx.loss3.data_ptr(), cx.dlogits.data_ptr(),
                   None, None, B, self.N, sp)
        return cx

    # ------------------------------------------------------------------ backward
    def backward(self, cx, feature_grad=None, pose=True):
        """Accumulates parameter gradients of the pass recorded in `cx` into store.grads.
        pose=True: backprop the keypoint loss (dlogits from forward).  feature_grad: extra
        gradient wrt base[17] output (DANN domain branch, already sign-reversed)."""
        B, sp, dt, st = cx.B, L.stream_ptr(), self.dtype, self.store
        h, w = cx.fh, cx.fw
        M = B * h * w
        dO17 = self._buf(cx.dO, 17, cx.O[17].shape)
        bi17 = self._bi('base.17.conv.3')
        if pose:
            # ---- head
            p = 'extras.3.conv'
            yp = cx.Y['xp3']
            g = self._buf(cx.G, 'xp3', yp.shape)
            L.call('b200sp_head_bwd', cx.dlogits.data_ptr(), C.byref(cx.head_vt), self._w('head.0.weight'), g.data_ptr(),
                   self._wg('head.0.weight'), self._wg('head.0.bias'), self._bnbwd(cx, self._bi(p + '.4'), yp, L.ACT_RELU), B, h * w * 1024, 1024, self.N, dt, sp)
            # ---- extras.3 pointwise + depthwise
            dcat = self._buf(cx.dO, 'cat', cx.O['cat'].shape)
            self._convdw_bwd(cx, 3, 1280, 1024, self._vt_plain(cx.O['cat']), dcat, None, None)
            # ---- RouterV2 backward: split dcat
            yr, yp1 = cx.Y['xr'], cx.Y['xp1']
            g_r, g_1 = self._buf(cx.G, 'xr', yr.shape), self._buf(cx.G, 'xp1', yp1.shape)
            bir, bi1 = self._bi('extras.2.conv.1'), self._bi('extras.1.conv.4')
            L.call('b200sp_reorg_cat_bwd', dcat.data_ptr(), g_r.data_ptr(), g_1.data_ptr(),
                   self._bnbwd(cx, bir, yr, L.ACT_LEAKY02), self._bnbwd(cx, bi1, yp1, L.ACT_RELU), B, h, w, 64, 1024, dt, sp)
            h13, w13 = cx.O[13].shape[1], cx.O[13].shape[2]
            dyr = self._vt_dy(cx, g_r, yr, bir)
            self._wgrad(C.byref(dyr), C.byref(self._vt_plain(cx.O[13])), self._wg('extras.2.conv.0.weight'), None,
                        B * h13 * w13, 64, 96, dt)
            d13p = self._buf(cx.dO, '13r', cx.O[13].shape)
            L.call('b200sp_pw_dgrad', C.byref(dyr), self._wq('extras.2.conv.0.weight'), None, 1.0, d13p.data_ptr(), None,
                   B * h13 * w13, 64, 96, self.gemm_dtype, sp)
            # ---- extras.1, extras.0
            vt_e0 = self._vt_bnact(cx, cx.Y['xp0'], self._bi('extras.0.conv.4'), L.ACT_RELU)
            g_e0p = self._buf(cx.G, 'xp0', cx.Y['xp0'].shape)
            self._convdw_bwd(cx, 1, 1024, 1024, vt_e0, g_e0p, self._bnbwd(cx, self._bi('extras.0.conv.4'), cx.Y['xp0'], L.ACT_RELU), None)
            self._convdw_bwd(cx, 0, 320, 1024, self._vt_plain(cx.O[17]), dO17,
                             self._bnbwd(cx, bi17, cx.Y['p17'], L.ACT_NONE), feature_grad)
        else:
            assert feature_grad is not None
            dO17.copy_(feature_grad.view(dO17.shape))
            L.call('b200sp_bn_bwd_reduce', dO17.data_ptr(), self._bnbwd(cx, bi17, cx.Y['p17'], L.ACT_NONE), M, 320, dt, sp)
        # ---- MobileNetV2 blocks 17..1
        for b in reversed(self.blocks):
            i, cin, cout, s, t = b['idx'], b['cin'], b['cout'], b['stride'], b['t']
            p = 'base.%d.conv' % i
            hid, j = cin * t, (0 if t == 1 else 1)
            hi_, wi_ = cx.geo[i]                                   # block input spatial size
            ho_, wo_ = (hi_ - 1) // s + 1, (wi_ - 1) // s + 1
            Mo, Mi = B * ho_ * wo_, B * hi_ * wi_
            bip, bid = self._bi('%s.%d' % (p, j + 2)), self._bi('%s.%d.1' % (p, j))
            yp, yd = cx.Y['p%d' % i], cx.Y['d%d' % i]
            dyp = self._vt_dy(cx, cx.dO[i], yp, bip)
            dvt = self._vt_bnact(cx, yd, bid, L.ACT_RELU6)
            self._wgrad(C.byref(dyp), C.byref(dvt), self._wg('%s.%d.weight' % (p, j + 1)), None, Mo, cout, hid, dt)
            gd = self._buf(cx.G, 'd%d' % i, yd.shape)
            L.call('b200sp_pw_dgrad', C.byref(dyp), self._wq('%s.%d.weight' % (p, j + 1)), None, 1.0, gd.data_ptr(),
                   self._bnbwd(cx, bid, yd, L.ACT_RELU6), Mo, cout, hid, self.gemm_dtype, sp)
            dyd = self._vt_dy(cx, gd, yd, bid)
            if t != 1:
                ye, bie = cx.Y['e%d' % i], self._bi(p + '.0.1')
                evt = self._vt_bnact(cx, ye, bie, L.ACT_RELU6)
                ge = self._buf(cx.G, 'e%d' % i, ye.shape)
                L.call('b200sp_dw_bwd', C.byref(dyd), C.byref(evt), self._w('%s.%d.0.weight' % (p, j)), None, ge.data_ptr(),
                       self._wg('%s.%d.0.weight' % (p, j)), self._bnbwd(cx, bie, ye, L.ACT_RELU6), B, hi_, wi_, hid, s, dt, sp)
                dye = self._vt_dy(cx, ge, ye, bie)
                xin = cx.O[i - 1]
                self._wgrad(C.byref(dye), C.byref(self._vt_plain(xin)), self._wg(p + '.0.0.weight'), None, Mi, hid, cin, dt)
                # gradient wrt the block input = dgrad (+ residual skip) (+ RouterV2 branch for base[13])
                skip = cx.dO[i] if b['res'] else (cx.dO['13r'] if (i == 14 and pose) else None)
                dprev = self._buf(cx.dO, i - 1, xin.shape)
                pprev = 'base.%d.conv' % (i - 1)
                tprev = self.blocks[i - 2]['t']
                biprev = self._bi('%s.%d' % (pprev, 2 if tprev == 1 else 3))
                L.call('b200sp_pw_dgrad', C.byref(dye), self._wq(p + '.0.0.weight'), skip.data_ptr() if skip is not None else None, 1.0,
                       dprev.data_ptr(), self._bnbwd(cx, biprev, cx.Y['p%d' % (i - 1)], L.ACT_NONE), Mi, hid, cin, self.gemm_dtype, sp)
            else:
                # block 1: depthwise reads the stem activation directly
                y0, bi0 = cx.Y['stem'], self._bi('base.0.1')
                svt = self._vt_bnact(cx, y0, bi0, L.ACT_RELU6)
                g0 = self._buf(cx.G, 'stem', y0.shape)
                L.call('b200sp_dw_bwd', C.byref(dyd), C.byref(svt), self._w('%s.0.0.weight' % p), None, g0.data_ptr(),
                       self._wg('%s.0.0.weight' % p), self._bnbwd(cx, bi0, y0, L.ACT_RELU6), B, hi_, wi_, hid, s, dt, sp)
                dy0 = self._vt_dy(cx, g0, y0, bi0)
                L.call('b200sp_stem_wgrad', cx.images.data_ptr(), C.byref(dy0), self._wg('base.0.0.weight'), B, cx.H, cx.W, dt, sp)
        self._join_side()

    # ------------------------------------------------------------------ DANN domain classifier
    def domain_forward(self, cx, label, loss_slot=0):
        """revgrad.py:75-80,92-93 + dann.py:85-92 on the base[17] output of the pass recorded in `cx`:
        GRL (identity forward) -> conv1x1 320->1280 (+bias, ReLU) -> AvgPool(7) -> conv1x1 1280->1 ->
        mean BCE-with-logits against the constant `label` (1 source, 0 target).  Fills cx.dom_z [B],
        cx.dom_loss [1] and the logit gradient cx.dom_dz."""
        assert self.dann
        B, sp, dt = cx.B, L.stream_ptr(), self.dtype
        h, w = cx.fh, cx.fw
        M = B * h * w
        f32 = dict(dtype=torch.float32, device=self.device)
        if getattr(cx, 'dom_z', None) is None or cx.dom_z.shape[0] != B:
            cx.dom_z, cx.dom_dz = torch.zeros(B, **f32), torch.zeros(B, **f32)
            cx.dom_pool, cx.dom_loss = torch.zeros(B, 1280, **f32), torch.zeros(1, **f32)
        hbuf = self._buf(cx.Y, 'dom_h', (B, h, w, 1280))
        fvt = self._vt_plain(cx.O[17])
        L.call('b200sp_pw_fwd', C.byref(fvt), self.store.w_ptr('domain_classifier.0.weight'),
               self.store.w_ptr('domain_classifier.0.bias'), L.ACT_RELU, hbuf.data_ptr(), None, M, 1280, 320, self.gemm_dtype, sp)
        L.call('b200sp_dann_head_fwd', hbuf.data_ptr(), self.store.w_ptr('domain_classifier.3.weight'),
               self.store.w_ptr('domain_classifier.3.bias'), cx.dom_pool.data_ptr(), cx.dom_z.data_ptr(), B, h * w, 1280, dt, sp)
        L.call('b200sp_bce_logits', cx.dom_z.data_ptr(), float(label), cx.dom_loss.data_ptr(), cx.dom_dz.data_ptr(), None, B, sp)
        return cx.dom_z

    def domain_backward(self, cx, neg_alpha_dev):
        """Backward of domain_forward: accumulates the domain-classifier gradients and returns the
        gradient entering base[17]'s output, already multiplied by -alpha (the gradient reversal,
        revgrad.py:52-56; `neg_alpha_dev` is a 1-element DEVICE tensor holding -alpha)."""
        B, sp, dt, st = cx.B, L.stream_ptr(), self.dtype, self.store
        h, w = cx.fh, cx.fw
        M = B * h * w
        hbuf = cx.Y['dom_h']
        L.call('b200sp_dann_head_bwd', hbuf.data_ptr(), cx.dom_dz.data_ptr(), cx.dom_pool.data_ptr(),
               st.w_ptr('domain_classifier.3.weight'), st.wg_ptr('domain_classifier.3.weight'),
               st.wg_ptr('domain_classifier.3.bias'), B, h * w, 1280, dt, sp)
        dvt = self._vt_plain(hbuf)                           # now holds dL/d(conv0 output)
        L.call('b200sp_pw_wgrad', C.byref(dvt), C.byref(self._vt_plain(cx.O[17])), st.wg_ptr('domain_classifier.0.weight'),
               st.wg_ptr('domain_classifier.0.bias'), M, 1280, 320, self.gemm_dtype, sp)
        fg = self._buf(cx.dO, 'dom_f', cx.O[17].shape)
        L.call('b200sp_pw_dgrad', C.byref(dvt), st.w_ptr('domain_classifier.0.weight'), None, 1.0, fg.data_ptr(), None,
               M, 1280, 320, self.gemm_dtype, sp)
        L.call('b200sp_scale_dev', fg.data_ptr(), fg.numel(), neg_alpha_dev.data_ptr(), 1.0, dt, sp)
        return fg

    def _convdw_bwd(self, cx, e, cin, cout, in_vt, g_in, in_bn, skip):
        """backward of extras.<e> ConvDw given g of its output BN already in cx.G['xp<e>'] (+coefs)."""
        B, h, w, sp, dt = cx.B, cx.fh, cx.fw, L.stream_ptr(), self.dtype
        M = B * h * w
        p = 'extras.%d.conv' % e
        yp, yd = cx.Y['xp%d' % e], cx.Y['xd%d' % e]
        bip, bid = self._bi(p + '.4'), self._bi(p + '.1')
        dyp = self._vt_dy(cx, cx.G['xp%d' % e], yp, bip)
        dvt = self._vt_bnact(cx, yd, bid, L.ACT_RELU)
        self._wgrad(C.byref(dyp), C.byref(dvt), self._wg(p + '.3.weight'), None, M, cout, cin, dt)
        gd = self._buf(cx.G, 'xd%d' % e, yd.shape)
        L.call('b200sp_pw_dgrad', C.byref(dyp), self._wq(p + '.3.weight'), None, 1.0, gd.data_ptr(),
               self._bnbwd(cx, bid, yd, L.ACT_RELU), M, cout, cin, self.gemm_dtype, sp)
        dyd = self._vt_dy(cx, gd, yd, bid)
        L.call('b200sp_dw_bwd', C.byref(dyd), C.byref(in_vt), self._w(p + '.0.weight'), skip.data_ptr() if skip is not None else None,
               g_in.data_ptr(), self._wg(p + '.0.weight'), in_bn, B, h, w, cin, 1, dt, sp)
