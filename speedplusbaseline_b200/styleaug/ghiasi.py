"""Ghiasi style-transfer network, forward only -- the layer plan, HBM buffers and launch sequence behind
StyleAugmentor.  Mirrors /root/reference/src/styleaug/ghiasi.py:106-135 (ConvInRelu :6-23,
UpsampleConvInRelu :26-62, ResidualBlock :65-103); every op is a libb200sp launch (include/b200sp.h).

Per convolution:   b200sp_convtc_fwd (TMA + tcgen05 shifted GEMM, bf16 operands, fp32 accumulate; raw fp32
output + InstanceNorm sums in its epilogue)  ->  b200sp_in_finalize (scale/shift per image & channel, folding
the style-conditional gamma/beta)  ->  b200sp_in_apply (normalise + affine + ReLU (+ residual), written
directly in the NEXT convolution's input layout: reflection padding, x2 nearest upsampling and the stride-2
phase split are destination-side index arithmetic).  The conv biases of the reference are dropped: each conv
feeds a mean-subtracting InstanceNorm, so they cancel exactly."""
import ctypes as C

import torch

from .. import _lib as L

EPS = 1e-5            # nn.InstanceNorm2d default


def cond_sets():
    """conditional-instance-norm parameter sets in packing order: (state_dict prefix, suffix, C)"""
    out = []
    for i in range(3, 8):
        out += [('layers.%d' % i, '1', 128), ('layers.%d' % i, '2', 128)]
    out += [('layers.8', '', 64), ('layers.9', '', 32), ('layers.10', '', 3)]
    return out


CONV_SPECS = ((0, 3, 32, 9), (1, 32, 64, 3), (2, 64, 128, 3), (8, 128, 64, 3), (9, 64, 32, 3), (10, 32, 3, 9))   # layer, Cin, Cout, k


def param_shapes():
    """state_dict key -> shape of the reference's `Ghiasi()` module (ghiasi.py:106-121; 84 tensors, the keys of
    checkpoint_transformer.pth['state_dict_ghiasi'])."""
    d = {}
    for i, ci, co, k in CONV_SPECS:
        d['layers.%d.conv.weight' % i], d['layers.%d.conv.bias' % i] = (co, ci, k, k), (co,)
    for i in range(3, 8):
        for j in ('1', '2'):
            d['layers.%d.conv%s.weight' % (i, j)], d['layers.%d.conv%s.bias' % (i, j)] = (128, 128, 3, 3), (128,)
    for pre, sfx, c in cond_sets():
        for g in ('beta', 'gamma'):
            d['%s.fc_%s%s.weight' % (pre, g, sfx)], d['%s.fc_%s%s.bias' % (pre, g, sfx)] = (c, 100), (c,)
    return d


def synthetic_state(seed=7):
    """random stand-in for the style checkpoints (benchmarks / smoke runs on a box without the reference's files):
    dict(ghiasi, mean, cov, base) as StyleAugmentor(state=...) takes it."""
    g = torch.Generator().manual_seed(seed)
    sd = {k: 0.05 * torch.randn(shp, generator=g) for k, shp in sorted(param_shapes().items())}
    cov = torch.randn(100, 100, generator=g)
    return dict(ghiasi=sd, mean=torch.randn(1, 100, generator=g), cov=(cov @ cov.t() / 100).numpy(), base=torch.randn(100, generator=g))


def plan_chunks(k, ps, Cin, Cp, Wq):
    """K-chunk list of the shifted GEMM.  Returns (chunks [(plane, c0, shift)], cols [[(kh, kw, c) | None] * 64])."""
    cbox = min(Cp, 64)
    P = 64 // cbox
    chunks, cols = [], []
    for kh in range(k):
        qy, dh = kh % ps, kh // ps
        for qx in range(ps):
            dws = {kw // ps: kw for kw in range(k) if kw % ps == qx}
            if not dws:
                continue
            for dw0 in range(0, max(dws) + 1, P):
                for cc in range(0, Cp, cbox):
                    ent = []
                    for p in range(P):
                        kw = dws.get(dw0 + p)
                        for c in range(cbox):
                            ent.append((kh, kw, cc + c) if (kw is not None and cc + c < Cin) else None)
                    chunks.append((qy * ps + qx, cc, dh * Wq + dw0))
                    cols.append(ent)
    return chunks, cols


def pack_weight(w, cols, N_pad):
    """OIHW fp32 conv weight -> bf16 [N_pad][64*n_chunks] in chunk order (zeros for padding taps/channels)."""
    Co = w.shape[0]
    K = 64 * len(cols)
    idx = torch.zeros(K, dtype=torch.long)
    mask = torch.zeros(K, dtype=torch.bool)
    kk = w.shape[2]
    for j, ent in enumerate(cols):
        for i, e in enumerate(ent):
            if e is not None:
                kh, kw, c = e
                idx[j * 64 + i] = (c * kk + kh) * kk + kw
                mask[j * 64 + i] = True
    flat = w.reshape(Co, -1).float().cpu()
    m = torch.zeros(N_pad, K, dtype=torch.float32)
    m[:Co] = flat[:, idx] * mask.float()[None, :]
    return m.to(torch.bfloat16).contiguous()


class _Conv:
    """one convolution: packed weights, chunk plan, descriptor (rebuilt when the geometry changes)."""

    def __init__(self, name, w, k, stride, Cp, device):
        self.name, self.k, self.ps, self.Cp = name, k, stride, Cp
        self.Co, self.Ci = w.shape[0], w.shape[1]
        self.N_pad = max(16, (self.Co + 15) // 16 * 16)
        self.N_out = (self.Co + 3) // 4 * 4
        self.w_ref = w
        self.device = device
        self._geo = None

    def setup(self, B, Hq, Wq, Ho, Wo, planes, out, stats):
        geo = (B, Hq, Wq, Ho, Wo)
        if self._geo != geo:
            chunks, cols = plan_chunks(self.k, self.ps, self.Ci, self.Cp, Wq)
            assert len(chunks) <= L.CONVTC_MAX_CHUNKS, (self.name, len(chunks))
            self.wp = pack_weight(self.w_ref, cols, self.N_pad).to(self.device)
            self.chunks = chunks
            self._geo = geo
        d = L.ConvDesc()
        for i in range(4):
            d.planes[i] = planes[i].data_ptr() if i < len(planes) else None
        d.w, d.out, d.stats = self.wp.data_ptr(), out.data_ptr(), (stats.data_ptr() if stats is not None else None)
        d.C, d.B, d.Hq, d.Wq, d.Ho, d.Wo = self.Cp, B, Hq, Wq, Ho, Wo
        d.N_pad, d.N_out, d.n_chunks = self.N_pad, self.N_out, len(self.chunks)
        for j, (pl, c0, sh) in enumerate(self.chunks):
            d.chunks[j].plane, d.chunks[j].c0, d.chunks[j].shift = pl, c0, sh
        self.desc = d
        return d


class _RowConv:
    """k x k stride-1 convolution with <= 4 output channels as a k-tap shifted GEMM producing T[m'][kw*Co + co]
    (see b200sp_conv_kwsum): the tensor-core pass has k taps instead of k*k and N = k*Co instead of Co."""

    def __init__(self, name, w, Cp, device):
        self.name, self.k, self.Cp, self.device = name, w.shape[2], Cp, device
        self.Co, self.Ci = w.shape[0], w.shape[1]
        self.Nt = (self.k * self.Co + 15) // 16 * 16          # MMA N
        self.N_pad = 16                                       # stats row width of the finished conv
        self.N_out = 4
        self.w_ref = w
        self._geo = None

    def setup(self, B, Hq, Wq, Ho, planes, T):
        geo = (B, Hq, Wq, Ho)
        k, Co, Ci, Cp = self.k, self.Co, self.Ci, self.Cp
        cbox = min(Cp, 64)
        P = 64 // cbox
        assert Cp <= 64
        if self._geo != geo:
            # chunk kh: plane rows m' + kh*Wq, first pixel of the 128-byte row only (the other P-1 pixels meet zero weights)
            m = torch.zeros(self.Nt, 64 * k, dtype=torch.float32)
            wr = self.w_ref.float().cpu()
            for kh in range(k):
                for kw in range(k):
                    for co in range(Co):
                        m[kw * Co + co, kh * 64:kh * 64 + Ci] = wr[co, :, kh, kw]
            self.wp = m.to(torch.bfloat16).contiguous().to(self.device)
            self.chunks = [(0, 0, kh * Wq) for kh in range(k)]
            self._geo = geo
        d = L.ConvDesc()
        for i in range(4):
            d.planes[i] = planes[i].data_ptr() if i < len(planes) else None
        d.w, d.out, d.stats = self.wp.data_ptr(), T.data_ptr(), None
        d.C, d.B, d.Hq, d.Wq, d.Ho, d.Wo = Cp, B, Hq, Wq, Ho, Wq      # every plane column of the first Ho rows is a valid T row
        d.N_pad, d.N_out, d.n_chunks = self.Nt, self.Nt, k
        for j, (pl, c0, sh) in enumerate(self.chunks):
            d.chunks[j].plane, d.chunks[j].c0, d.chunks[j].shift = pl, c0, sh
        return d


class _UpConv:
    """nearest x2 upsample + ReflectionPad2d(1) + 3x3 conv (ghiasi.py:34-37) in SUB-PIXEL form: output pixel (2i+a, 2j+b)
    only ever sees source pixels i-1+a .. i+a (rows) and j-1+b .. j+b (columns), so the layer is four 2x2 convolutions
    on the source-resolution plane with the 3x3 taps that land on the same source pixel summed.  Reflection of the
    upsampled image at its border is edge replication in source coordinates (upsampled pixel -1 -> 1 = source 0), so the
    plane is the source image replicate-padded by one pixel.  2.25x fewer MACs, and no 4x-size upsampled plane exists."""
    TAPS = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}      # (phase, plane offset) -> 3x3 tap indices

    def __init__(self, name, w, device):
        self.name, self.device = name, device
        self.Co, self.Ci = w.shape[0], w.shape[1]
        assert self.Ci % 64 == 0
        self.N_pad = max(16, (self.Co + 15) // 16 * 16)
        self.N_out = (self.Co + 3) // 4 * 4
        self.w_ref = w.float().cpu()
        self._geo = None

    def setup(self, B, Hs, Ws, plane, out, stats):
        """plane: replicate-padded source [B][Hs+2][Ws+2][Ci]; out: [B][2Hs][2Ws][N_out]; returns the 4 phase descriptors."""
        Hq, Wq = Hs + 2, Ws + 2
        if self._geo != (B, Hs, Ws):
            self.wp, self.chunks = [], []
            for a in range(2):
                for b in range(2):
                    cols, ch = [], []
                    for ty in range(2):
                        for tx in range(2):
                            wsum = sum(self.w_ref[:, :, kh, kw] for kh in self.TAPS[(a, ty)] for kw in self.TAPS[(b, tx)])    # [Co][Ci]
                            for cc in range(0, self.Ci, 64):
                                cols.append(wsum[:, cc:cc + 64])
                                ch.append((0, cc, (a + ty) * Wq + (b + tx)))
                    m = torch.zeros(self.N_pad, 64 * len(cols))
                    for j, blk in enumerate(cols):
                        m[:self.Co, j * 64:(j + 1) * 64] = blk
                    self.wp.append(m.to(torch.bfloat16).contiguous().to(self.device))
                    self.chunks.append(ch)
            self._geo = (B, Hs, Ws)
        descs = []
        for a in range(2):
            for b in range(2):
                q = a * 2 + b
                d = L.ConvDesc()
                d.planes[0] = plane.data_ptr()
                d.w, d.out, d.stats = self.wp[q].data_ptr(), out.data_ptr(), stats.data_ptr()
                d.C, d.B, d.Hq, d.Wq, d.Ho, d.Wo = self.Ci, B, Hq, Wq, Hs, Ws
                d.N_pad, d.N_out, d.n_chunks = self.N_pad, self.N_out, len(self.chunks[q])
                for j, (pl, c0, sh) in enumerate(self.chunks[q]):
                    d.chunks[j].plane, d.chunks[j].c0, d.chunks[j].shift = pl, c0, sh
                d.OH, d.OW, d.sy, d.sx, d.oy, d.ox = 2 * Hs, 2 * Ws, 2, 2, a, b
                descs.append(d)
        return descs


class GhiasiEngine:
    def __init__(self, state_dict, device):
        L.require_cuda()
        self.device = torch.device(device)
        sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        dev = self.device
        self.convs = {}

        def add(name, key, k, stride, Cp):
            self.convs[name] = _Conv(name, sd[key], k, stride, Cp, dev)

        add('c0', 'layers.0.conv.weight', 9, 1, 8)
        add('c1', 'layers.1.conv.weight', 3, 2, 32)
        add('c2', 'layers.2.conv.weight', 3, 2, 64)
        for i in range(3, 8):
            add('r%da' % i, 'layers.%d.conv1.weight' % i, 3, 1, 128)
            add('r%db' % i, 'layers.%d.conv2.weight' % i, 3, 1, 128)
        add('c8', 'layers.8.conv.weight', 3, 1, 128)
        add('c9', 'layers.9.conv.weight', 3, 1, 64)
        add('c10', 'layers.10.conv.weight', 9, 1, 32)
        self.c10_row = _RowConv('c10row', sd['layers.10.conv.weight'], 32, dev)
        self.row_decomposed_c10 = True
        self.up8, self.up9 = _UpConv('c8up', sd['layers.8.conv.weight'], dev), _UpConv('c9up', sd['layers.9.conv.weight'], dev)
        self.subpixel_upconvs = True
        # the 26 Linear(100 -> C) of the conditional instance norms, concatenated: per set [gamma | beta]
        Ws, bs, self.gb_off, off = [], [], {}, 0
        for p, sfx, Cc in cond_sets():
            for gname in ('gamma', 'beta'):
                Ws.append(sd['%s.fc_%s%s.weight' % (p, gname, sfx)])
                bs.append(sd['%s.fc_%s%s.bias' % (p, gname, sfx)])
            self.gb_off[(p, sfx)] = (off, off + Cc)
            off += 2 * Cc
        self.T = off
        self.Wcat = torch.cat(Ws, 0).contiguous().to(dev)            # [T][100]
        self.bcat = torch.cat(bs, 0).contiguous().to(dev)
        self._bufs = {}
        self._keep = []

    # ------------------------------------------------------------------ buffers
    def _buf(self, name, shape, dtype, zero=False):
        t = self._bufs.get(name)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.zeros(n, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t[:n].view(shape)

    def _plane(self, name, B, Hd, Wd, Cd):
        # 16 pixels of slack: narrow-channel planes are read a few pixels past the last row (zero weights)
        t = self._buf(name, (B * Hd * Wd + 16, Cd), torch.bfloat16)
        return t

    # ------------------------------------------------------------------ one conv + norm
    def _conv(self, name, planes, B, Hq, Wq, Ho, Wo):
        cv = self.convs[name]
        raw = self._buf('raw_%dx%dx%d' % (Ho, Wo, cv.N_out), (B, Ho, Wo, cv.N_out), torch.float32)
        stats = self._buf('stats_%d' % cv.N_pad, (B, 2, cv.N_pad), torch.float32)
        d = cv.setup(B, Hq, Wq, Ho, Wo, planes, raw, stats)
        self._keep.append(d)
        L.call('b200sp_convtc_fwd', C.byref(d), L.stream_ptr())
        return raw, stats, cv

    def _upconv(self, uc, plane, B, Hs, Ws):
        raw = self._buf('raw_%dx%dx%d' % (2 * Hs, 2 * Ws, uc.N_out), (B, 2 * Hs, 2 * Ws, uc.N_out), torch.float32)
        stats = self._buf('stats_%d' % uc.N_pad, (B, 2, uc.N_pad), torch.float32)
        for d in uc.setup(B, Hs, Ws, plane, raw, stats):
            self._keep.append(d)
            L.call('b200sp_convtc_fwd', C.byref(d), L.stream_ptr())
        return raw, stats, uc

    def _finalize(self, stats, cv, B, HW, gb, key):
        Cc = cv.Co
        scale = self._buf('scale', (B, 128), torch.float32).view(-1)[:B * Cc]
        shift = self._buf('shift', (B, 128), torch.float32).view(-1)[:B * Cc]
        if key is None:
            g = b = None
        else:
            og, ob = self.gb_off[key]
            g, b = gb.data_ptr() + 4 * og, gb.data_ptr() + 4 * ob
        L.call('b200sp_in_finalize', stats.data_ptr(), g, b, self.T, scale.data_ptr(), shift.data_ptr(), B, Cc, cv.N_pad, HW, EPS,
               L.stream_ptr())
        return scale, shift

    def _apply(self, raw, scale, shift, B, Hs, Ws, Cc, act, pad, up, ps, dst_name, Cd=None, res_in=None, res_out=None, pad_mode=0):
        Cd = Cd or Cc
        Hd, Wd = (Hs * up + 2 * pad) // ps, (Ws * up + 2 * pad) // ps
        planes = [self._plane('%s_%d' % (dst_name, q), B, Hd, Wd, Cd) for q in range(ps * ps)]
        d = L.InApplyDesc()
        d.raw, d.scale, d.shift = raw.data_ptr(), scale.data_ptr(), shift.data_ptr()
        d.res_in = res_in.data_ptr() if res_in is not None else None
        d.res_out = res_out.data_ptr() if res_out is not None else None
        for q in range(4):
            d.planes[q] = planes[q].data_ptr() if q < len(planes) else None
        d.B, d.Hs, d.Ws, d.Cs, d.C, d.act = B, Hs, Ws, raw.shape[-1], Cc, act
        d.pad, d.up, d.ps, d.Hd, d.Wd, d.Cd = pad, up, ps, Hd, Wd, Cd
        d.pad_mode = pad_mode
        self._keep.append(d)
        L.call('b200sp_in_apply', C.byref(d), L.stream_ptr())
        return planes, Hd, Wd

    # ------------------------------------------------------------------ forward
    def forward(self, x, embedding, out=None):
        """x [B,3,H,W] fp32 NCHW in [0,1] on device, embedding [B,100] fp32 on device -> [B,3,H,W] in (0,1)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B, _, H, W = x.shape
        assert H % 4 == 0 and W % 4 == 0 and H >= 16 and W >= 16, 'style net needs H, W multiples of 4'
        sp = L.stream_ptr()
        self._keep.clear()
        gb = self._buf('gb', (B, self.T), torch.float32)
        L.call('b200sp_style_linear', embedding.data_ptr(), self.Wcat.data_ptr(), self.bcat.data_ptr(), gb.data_ptr(), B, 100, self.T, sp)
        # layer 0: reflpad 4 + conv 9x9 3->32 + IN + ReLU
        p0 = self._plane('p0', B, H + 8, W + 8, 8)
        L.call('b200sp_sa_prep', x.data_ptr(), p0.data_ptr(), B, H, W, 4, 8, sp)
        raw, st, cv = self._conv('c0', [p0], B, H + 8, W + 8, H, W)
        sc, sh = self._finalize(st, cv, B, H * W, gb, None)
        pl, Hd, Wd = self._apply(raw, sc, sh, B, H, W, 32, L.ACT_RELU, 1, 1, 2, 'p1')
        # layer 1: 3x3 s2 32->64
        H1, W1 = H // 2, W // 2
        raw, st, cv = self._conv('c1', pl, B, Hd, Wd, H1, W1)
        sc, sh = self._finalize(st, cv, B, H1 * W1, gb, None)
        pl, Hd, Wd = self._apply(raw, sc, sh, B, H1, W1, 64, L.ACT_RELU, 1, 1, 2, 'p2')
        # layer 2: 3x3 s2 64->128; its output is the first residual-stream tensor
        H2, W2 = H1 // 2, W1 // 2
        raw, st, cv = self._conv('c2', pl, B, Hd, Wd, H2, W2)
        sc, sh = self._finalize(st, cv, B, H2 * W2, gb, None)
        rs = [self._buf('rs0', (B, H2, W2, 128), torch.float32), self._buf('rs1', (B, H2, W2, 128), torch.float32)]
        pl, Hd, Wd = self._apply(raw, sc, sh, B, H2, W2, 128, L.ACT_RELU, 1, 1, 1, 'pa', res_out=rs[0])
        cur = 0
        for i in range(3, 8):
            p = 'layers.%d' % i
            raw, st, cv = self._conv('r%da' % i, pl, B, Hd, Wd, H2, W2)
            sc, sh = self._finalize(st, cv, B, H2 * W2, gb, (p, '1'))
            plb, _, _ = self._apply(raw, sc, sh, B, H2, W2, 128, L.ACT_RELU, 1, 1, 1, 'pb')
            raw, st, cv = self._conv('r%db' % i, plb, B, Hd, Wd, H2, W2)
            sc, sh = self._finalize(st, cv, B, H2 * W2, gb, (p, '2'))
            if i < 7:
                pl, Hd, Wd = self._apply(raw, sc, sh, B, H2, W2, 128, L.ACT_NONE, 1, 1, 1, 'pa', res_in=rs[cur], res_out=rs[cur ^ 1])
                cur ^= 1
            elif self.subpixel_upconvs:   # block 7 feeds the first upsampling conv (sub-pixel form: replicate-padded source plane)
                pl, Hd, Wd = self._apply(raw, sc, sh, B, H2, W2, 128, L.ACT_NONE, 1, 1, 1, 'pu8', res_in=rs[cur], pad_mode=1)
            else:   # block 7 feeds the first upsampling conv: x2 nearest + reflpad 1
                pl, Hd, Wd = self._apply(raw, sc, sh, B, H2, W2, 128, L.ACT_NONE, 1, 2, 1, 'pu8', res_in=rs[cur])
        # layer 8: up x2 + 3x3 128->64, cond IN, ReLU
        if self.subpixel_upconvs:
            raw, st, cv = self._upconv(self.up8, pl[0], B, H2, W2)
            sc, sh = self._finalize(st, cv, B, H1 * W1, gb, ('layers.8', ''))
            pl, Hd, Wd = self._apply(raw, sc, sh, B, H1, W1, 64, L.ACT_RELU, 1, 1, 1, 'pu9', pad_mode=1)
            raw, st, cv = self._upconv(self.up9, pl[0], B, H1, W1)
        else:
            raw, st, cv = self._conv('c8', pl, B, Hd, Wd, H1, W1)
            sc, sh = self._finalize(st, cv, B, H1 * W1, gb, ('layers.8', ''))
            pl, Hd, Wd = self._apply(raw, sc, sh, B, H1, W1, 64, L.ACT_RELU, 1, 2, 1, 'pu9')
            # layer 9: up x2 + 3x3 64->32, cond IN, ReLU; next conv is 9x9 -> reflpad 4
            raw, st, cv = self._conv('c9', pl, B, Hd, Wd, H, W)
        sc, sh = self._finalize(st, cv, B, H * W, gb, ('layers.9', ''))
        pl, Hd, Wd = self._apply(raw, sc, sh, B, H, W, 32, L.ACT_RELU, 4, 1, 1, 'p10')
        # layer 10: 9x9 32->3, cond IN, sigmoid
        if self.row_decomposed_c10:
            rc = self.c10_row
            T = self._buf('T10', (B, H, Wd, rc.Nt), torch.float32)
            d = rc.setup(B, Hd, Wd, H, pl, T)
            self._keep.append(d)
            L.call('b200sp_convtc_fwd', C.byref(d), sp)
            raw = self._buf('raw_%dx%dx%d' % (H, W, rc.N_out), (B, H, W, rc.N_out), torch.float32)
            st = self._buf('stats_%d' % rc.N_pad, (B, 2, rc.N_pad), torch.float32)
            L.call('b200sp_conv_kwsum', T.data_ptr(), raw.data_ptr(), st.data_ptr(), B, H, W, Wd, rc.Nt, rc.k, rc.Co, rc.N_out, rc.N_pad, sp)
            cv = self.convs['c10']
        else:
            raw, st, cv = self._conv('c10', pl, B, Hd, Wd, H, W)
        sc, sh = self._finalize(st, cv, B, H * W, gb, ('layers.10', ''))
        if out is None:
            out = torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device)
        L.call('b200sp_in_apply_final', raw.data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(), B, H, W, raw.shape[-1], 3, sp)
        return out
