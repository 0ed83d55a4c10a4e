"""SPN (AlexNet backbone + two 3-layer FC branches) execution engine: layer plan, HBM buffers, kernel sequencing.

Mirrors /root/reference/src/nets/spn.py:50-143 and the loss of src/core/trainer.py:152-165.  Every op is a
libb200sp launch: convolutions = b200sp_im2col + one tcgen05 GEMM per group with bias+ReLU in the epilogue
(3xTF32, fp32 patch matrix: the attitude-class argmax must match the fp32 reference bit for bit), pool+LRN fused,
FC layers = the same GEMM family (the 150 M FC weights make this path weight-bandwidth bound), soft-target
cross entropy with the gradient in the same pass.  NHWC activations; parameters in one flat store so
clip_grad_value_ + AdamW are one launch (optim.FusedAdamW, clip_mode=2)."""
import ctypes as C

import torch

from . import _lib as L
from .params import ParamStore

LRN_ALPHA, LRN_BETA = 2e-5, 0.75            # spn.py:62,67  LocalResponseNorm(2, alpha=2e-5, beta=0.75, k=1.0)

# name, Cin, Cout, k, stride, pad, groups
CONVS = [('conv1', 3, 96, 11, 4, 0, 1), ('conv2', 96, 256, 5, 1, 2, 2), ('conv3', 256, 384, 3, 1, 1, 1),
         ('conv4', 384, 384, 3, 1, 1, 2), ('conv5', 384, 256, 3, 1, 1, 2)]


def spn_layout(num_classes):
    W = []
    for name, ci, co, k, s, p, g in CONVS:
        W.append((name + '.weight', 'ohwi_pad4' if (k * k * ci // g) % 4 else 'ohwi', (co, ci // g, k, k)))
        W.append((name + '.bias', 'plain', (co,)))
    for name, i, o, kind in (('fc6', 9216, 4096, 'fc_chw:256x6x6'), ('fc7', 4096, 4096, 'plain'), ('fc8', 4096, num_classes, 'plain'),
                             ('fc9', 9216, 4096, 'fc_chw:256x6x6'), ('fc10', 4096, 4096, 'plain'), ('fc11', 4096, num_classes, 'plain')):
        W.append((name + '.weight', kind, (o, i)))
        W.append((name + '.bias', 'plain', (o,)))
    return W, [k for k, _, _ in W]


def _pad4(n):
    return (n + 3) // 4 * 4


class SPNEngine:
    def __init__(self, num_classes=5000, device=None):
        L.require_cuda()
        assert num_classes % 4 == 0, 'num_classes must be a multiple of 4 (16-byte GEMM granularity)'
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.nc = num_classes
        W, self.key_order = spn_layout(num_classes)
        self.store = ParamStore(W, [], self.device)
        self._bufs = {}
        self.keep = []
        self.drop_p = 0.5
        self.base_seed = 0                 # f(cfg.seed, rank): set by the training step / CLI
        self.drop_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)     # training forwards taken (device-resident)

    def _buf(self, name, shape, dtype=torch.float32):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    @staticmethod
    def _vt(ptr):
        return L.VTensor(ptr, None, None, None, None, L.VT_PLAIN, 0)

    def _mask_bn(self, y):
        """b200sp_bnbwd carrying only the ReLU mask of the saved activation `y` (no statistics)."""
        s = L.BnBwd(None, None, None, y.data_ptr(), None, None, None, None, None, None, None, None, None, L.ACT_RELU, 0)
        self.keep.append(s)
        return C.byref(s)

    # ------------------------------------------------------------------ convolution = im2col + GEMM per group
    def _conv_fwd(self, spec, x, B, H, W, nchw=False):
        name, ci, co, k, s, p, g = spec
        st, sp = self.store, L.stream_ptr()
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        Ig, Og, M = ci // g, co // g, B * Ho * Wo
        Kp = _pad4(k * k * Ig)
        y = self._buf('a_' + name, (B, Ho, Wo, co))
        for gi in range(g):
            col = self._buf('col_%s_%d' % (name, gi), (M, Kp))
            L.call('b200sp_im2col', x.data_ptr(), col.data_ptr(), B, H, W, ci, gi * Ig, Ig, k, s, p, Kp, 1 if nchw else 0, sp)
            vt = self._vt(col.data_ptr())
            self.keep.append(vt)
            L.call('b200sp_gemm_fwd', C.byref(vt), Kp, st.w_ptr(name + '.weight') + 4 * gi * Og * Kp, st.w_ptr(name + '.bias') + 4 * gi * Og,
                   L.ACT_RELU, y.data_ptr() + 4 * gi * Og, co, M, Og, Kp, L.F32, sp)
        return y, Ho, Wo

    def _conv_bwd(self, spec, dy, x_shape, dx, act_mask):
        """dy: gradient wrt the conv's pre-activation output (already ReLU-masked) [B,Ho,Wo,co]."""
        name, ci, co, k, s, p, g = spec
        B, H, W = x_shape
        st, sp = self.store, L.stream_ptr()
        Ho, Wo = dy.shape[1], dy.shape[2]
        Ig, Og, M = ci // g, co // g, B * Ho * Wo
        Kp = _pad4(k * k * Ig)
        vfull = self._vt(dy.data_ptr())
        self.keep.append(vfull)
        L.call('b200sp_colsum_f32', C.byref(vfull), st.wg_ptr(name + '.bias'), M, co, L.F32, sp)
        for gi in range(g):
            col = self._bufs['col_%s_%d' % (name, gi)]
            vdy, vcol = self._vt(dy.data_ptr() + 4 * gi * Og), self._vt(col.data_ptr())
            self.keep += [vdy, vcol]
            L.call('b200sp_gemm_wgrad', C.byref(vdy), co, C.byref(vcol), Kp, st.wg_ptr(name + '.weight') + 4 * gi * Og * Kp, M, Og, Kp, L.F32, sp)
            if dx is not None:
                dcol = self._buf('dcol', (self._dcol_elems,)).view(-1)[:M * Kp].view(M, Kp)
                L.call('b200sp_gemm_dgrad', C.byref(vdy), co, st.w_ptr(name + '.weight') + 4 * gi * Og * Kp, None, 1.0, dcol.data_ptr(), None,
                       M, Og, Kp, L.F32, sp)
                L.call('b200sp_col2im', dcol.data_ptr(), dx.data_ptr(), act_mask.data_ptr() if act_mask is not None else None,
                       B, H, W, ci, gi * Ig, Ig, k, s, p, Kp, sp)

    # ------------------------------------------------------------------ FC
    def _fc_fwd(self, name, x, B, K, N, act, tag):
        y = self._buf('h_' + tag, (B, N))
        sp = L.stream_ptr()
        if B <= 128:        # one M tile: split the K loop across CTAs (fp32 red.add), bias + ReLU in a tiny second pass
            y.zero_()
            L.call('b200sp_fc_fwd_splitk', x.data_ptr(), self.store.w_ptr(name + '.weight'), y.data_ptr(), B, N, K, sp)
            L.call('b200sp_bias_act', y.data_ptr(), self.store.w_ptr(name + '.bias'), B, N, 1 if act == L.ACT_RELU else 0, sp)
            return y
        vt = self._vt(x.data_ptr())
        self.keep.append(vt)
        L.call('b200sp_pw_fwd', C.byref(vt), self.store.w_ptr(name + '.weight'), self.store.w_ptr(name + '.bias'), act, y.data_ptr(), None,
               B, N, K, L.F32, sp)
        return y

    def _fc_bwd(self, name, dy, x, B, K, N, dx, skip, mask_y):
        st, sp = self.store, L.stream_ptr()
        vdy, vx = self._vt(dy.data_ptr()), self._vt(x.data_ptr())
        self.keep += [vdy, vx]
        L.call('b200sp_pw_wgrad', C.byref(vdy), C.byref(vx), st.wg_ptr(name + '.weight'), st.wg_ptr(name + '.bias'), B, N, K, L.F32, sp)
        if B <= 128:
            if skip is None:
                dx.zero_()
            else:
                assert skip.data_ptr() == dx.data_ptr()          # accumulate onto the other branch's gradient in place
            L.call('b200sp_fc_dgrad_splitk', dy.data_ptr(), st.w_ptr(name + '.weight'), dx.data_ptr(), B, N, K, sp)
            if mask_y is not None:
                L.call('b200sp_relu_mask', dx.data_ptr(), mask_y.data_ptr(), dx.numel(), sp)
            return
        L.call('b200sp_pw_dgrad', C.byref(vdy), st.w_ptr(name + '.weight'), skip.data_ptr() if skip is not None else None, 1.0, dx.data_ptr(),
               self._mask_bn(mask_y) if mask_y is not None else None, B, N, K, L.F32, sp)

    # ------------------------------------------------------------------ forward
    def forward(self, images, y_classes=None, y_weights=None, train=True):
        """images [B,3,227,227] fp32 NCHW on device.  Returns (c, r) logits [B, num_classes]; with targets also fills
        self.loss2 = (loss_class, loss_regress) and the logit gradients for backward()."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        B, _, H, W = images.shape
        sp = L.stream_ptr()
        self.keep.clear()
        self.B, self.HW = B, (H, W)
        self._dcol_elems = 0
        a1, H1, W1 = self._conv_fwd(CONVS[0], images, B, H, W, nchw=True)
        Hp1, Wp1 = (H1 - 3) // 2 + 1, (W1 - 3) // 2 + 1
        p1, n1 = self._buf('p1', (B, Hp1, Wp1, 96)), self._buf('n1', (B, Hp1, Wp1, 96))
        am1 = self._buf('am1', (B, Hp1, Wp1, 96), torch.uint8)
        L.call('b200sp_pool_lrn_fwd', a1.data_ptr(), p1.data_ptr(), n1.data_ptr(), am1.data_ptr(), B, H1, W1, 96, 1, LRN_ALPHA, LRN_BETA, sp)
        a2, H2, W2 = self._conv_fwd(CONVS[1], n1, B, Hp1, Wp1)
        Hp2, Wp2 = (H2 - 3) // 2 + 1, (W2 - 3) // 2 + 1
        p2, n2 = self._buf('p2', (B, Hp2, Wp2, 256)), self._buf('n2', (B, Hp2, Wp2, 256))
        am2 = self._buf('am2', (B, Hp2, Wp2, 256), torch.uint8)
        L.call('b200sp_pool_lrn_fwd', a2.data_ptr(), p2.data_ptr(), n2.data_ptr(), am2.data_ptr(), B, H2, W2, 256, 1, LRN_ALPHA, LRN_BETA, sp)
        a3, H3, W3 = self._conv_fwd(CONVS[2], n2, B, Hp2, Wp2)
        a4, _, _ = self._conv_fwd(CONVS[3], a3, B, H3, W3)
        a5, _, _ = self._conv_fwd(CONVS[4], a4, B, H3, W3)
        Hp5, Wp5 = (H3 - 3) // 2 + 1, (W3 - 3) // 2 + 1
        assert Hp5 * Wp5 * 256 == 9216, 'SPN needs 227x227 inputs (spn.py:80: 6*6*256 features)'
        f = self._buf('f', (B, Hp5, Wp5, 256))
        am5 = self._buf('am5', (B, Hp5, Wp5, 256), torch.uint8)
        L.call('b200sp_pool_lrn_fwd', a5.data_ptr(), None, f.data_ptr(), am5.data_ptr(), B, H3, W3, 256, 0, 0.0, 0.0, sp)
        self.geo = dict(H1=H1, W1=W1, Hp1=Hp1, Wp1=Wp1, H2=H2, W2=W2, Hp2=Hp2, Wp2=Wp2, H3=H3, W3=W3)
        self._dcol_elems = max(B * H2 * W2 * 1200, B * H3 * W3 * 2304)
        drop = train and self.drop_p > 0
        self.drop = drop
        if drop:               # one tick per training forward: every forward (eager, module facade or graph replay) draws new masks
            L.call('b200sp_add_i64', self.drop_ctr.data_ptr(), 1, 1, L.stream_ptr())
        out = []
        for bi, (fa, fb, fc) in enumerate((('fc6', 'fc7', 'fc8'), ('fc9', 'fc10', 'fc11'))):
            h1 = self._fc_fwd(fa, f, B, 9216, 4096, L.ACT_RELU, fa)
            h1d = self._dropout(h1, fa, bi * 2) if drop else h1
            h2 = self._fc_fwd(fb, h1d, B, 4096, 4096, L.ACT_RELU, fb)
            h2d = self._dropout(h2, fb, bi * 2 + 1) if drop else h2
            out.append(self._fc_fwd(fc, h2d, B, 4096, self.nc, L.ACT_NONE, fc))
        c, r = out
        self.has_loss = y_classes is not None
        if self.has_loss:
            rows = self._buf('loss_rows', (2, B))
            self.loss2 = self._buf('loss2', (2,))
            dzc, dzr = self._buf('dz_c', (B, self.nc)), self._buf('dz_r', (B, self.nc))
            # trainer.py:152-158: loss = CE(classes, yClasses) + 10 * CE(weights, yWeights)
            L.call('b200sp_soft_ce', c.data_ptr(), y_classes.data_ptr(), rows[0].data_ptr(), dzc.data_ptr(), B, self.nc, 1.0, sp)
            L.call('b200sp_soft_ce', r.data_ptr(), y_weights.data_ptr(), rows[1].data_ptr(), dzr.data_ptr(), B, self.nc, 10.0, sp)
            L.call('b200sp_soft_ce_mean', rows[0].data_ptr(), rows[1].data_ptr(), self.loss2.data_ptr(), B, sp)
        return c, r

    def _dropout(self, h, tag, idx):
        out = self._buf('hd_' + tag, tuple(h.shape))
        mask = self._buf('m_' + tag, tuple(h.shape), torch.uint8)
        seed = ((self.base_seed * 4 + idx + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        L.call('b200sp_dropout_fwd_ctr', h.data_ptr(), out.data_ptr(), mask.data_ptr(), h.numel(), float(self.drop_p), seed,
               self.drop_ctr.data_ptr(), L.stream_ptr())
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, scale_c=None, scale_r=None):
        """accumulates all parameter gradients of the last forward(with targets) into store.grads."""
        B, sp, bf, g = self.B, L.stream_ptr(), self._bufs, self.geo
        f = bf['f']
        df = self._buf('df', tuple(f.shape))
        first = True
        for (fa, fb, fc), dz in ((('fc6', 'fc7', 'fc8'), bf['dz_c']), (('fc9', 'fc10', 'fc11'), bf['dz_r'])):
            h1, h2 = bf['h_' + fa], bf['h_' + fb]
            h1d = bf['hd_' + fa] if self.drop else h1
            h2d = bf['hd_' + fb] if self.drop else h2
            dh2 = self._buf('dh_' + fb, tuple(h2.shape))
            self._fc_bwd(fc, dz, h2d, B, 4096, self.nc, dh2, None, h2)
            if self.drop:
                L.call('b200sp_dropout_bwd', dh2.data_ptr(), bf['m_' + fb].data_ptr(), dh2.numel(), float(self.drop_p), sp)
            dh1 = self._buf('dh_' + fa, tuple(h1.shape))
            self._fc_bwd(fb, dh2, h1d, B, 4096, 4096, dh1, None, h1)
            if self.drop:
                L.call('b200sp_dropout_bwd', dh1.data_ptr(), bf['m_' + fa].data_ptr(), dh1.numel(), float(self.drop_p), sp)
            self._fc_bwd(fa, dh1, f, B, 9216, 4096, df, None if first else df, None)
            first = False
        a1, a2, a3, a4, a5 = (bf['a_conv%d' % i] for i in range(1, 6))
        H3, W3 = g['H3'], g['W3']
        scratch = self._buf('lrn_scratch', (max(bf['p1'].numel(), bf['p2'].numel()),))
        da5 = self._buf('da5', tuple(a5.shape))
        L.call('b200sp_pool_lrn_bwd', df.data_ptr(), None, a5.data_ptr(), bf['am5'].data_ptr(), None, da5.data_ptr(), B, H3, W3, 256, 0, 0.0, 0.0, 1, sp)
        da4 = self._buf('da4', tuple(a4.shape))
        self._conv_bwd(CONVS[4], da5, (B, H3, W3), da4, a4)
        da3 = self._buf('da3', tuple(a3.shape))
        self._conv_bwd(CONVS[3], da4, (B, H3, W3), da3, a3)
        dn2 = self._buf('dn2', tuple(bf['n2'].shape))
        self._conv_bwd(CONVS[2], da3, (B, g['Hp2'], g['Wp2']), dn2, None)
        da2 = self._buf('da2', tuple(a2.shape))
        L.call('b200sp_pool_lrn_bwd', dn2.data_ptr(), bf['p2'].data_ptr(), a2.data_ptr(), bf['am2'].data_ptr(), scratch.data_ptr(), da2.data_ptr(),
               B, g['H2'], g['W2'], 256, 1, LRN_ALPHA, LRN_BETA, 1, sp)
        dn1 = self._buf('dn1', tuple(bf['n1'].shape))
        self._conv_bwd(CONVS[1], da2, (B, g['Hp1'], g['Wp1']), dn1, None)
        da1 = self._buf('da1', tuple(a1.shape))
        L.call('b200sp_pool_lrn_bwd', dn1.data_ptr(), bf['p1'].data_ptr(), a1.data_ptr(), bf['am1'].data_ptr(), scratch.data_ptr(), da1.data_ptr(),
               B, g['H1'], g['W1'], 96, 1, LRN_ALPHA, LRN_BETA, 1, sp)
        self._conv_bwd(CONVS[0], da1, (B,) + self.HW, None, None)
