"""Flat parameter / gradient / buffer store shared by the engines and the fused optimizer.

HBM layout (all fp32 unless noted, every segment 16-byte aligned):

    params : [ weights & biases, reference state_dict order, kernel-native layout | all BN gammas | all BN betas ]
    grads  : same layout (zeroed once per step; kernels accumulate)
    bufs   : [ all BN running_mean | all BN running_var ]      nbt : int64 [num BN]

Kernel-native layouts differ from the reference's OIHW only where the NHWC kernels need it
(depthwise [C,1,3,3] -> [9][C]; KRN head [N,C,7,7] -> [N][7][7][C]; SPN convs -> [O][kh][kw][I]);
`state_dict()` / `load_state_dict()` convert, so checkpoints written by the reference
(src/utils/utils.py:109-135) round-trip with identical keys and shapes.
"""
from collections import OrderedDict

import torch


def _align4(n):
    return (n + 3) // 4 * 4


def _align8(n):       # segment starts: 16-byte aligned in fp32 AND in the bf16 mirror
    return (n + 7) // 8 * 8


class Entry:
    __slots__ = ('key', 'kind', 'ref_shape', 'numel', 'off')

    def __init__(self, key, kind, ref_shape):
        self.key, self.kind, self.ref_shape = key, kind, tuple(ref_shape)
        self.numel = native_numel(kind, self.ref_shape)      # elements in the kernel-native layout
        self.off = -1


def native_numel(kind, ref_shape):
    n = 1
    for s in ref_shape:
        n *= s
    if kind == 'ohwi_pad4':                # rows padded to a multiple of 4 (16-byte GEMM granularity)
        k = n // ref_shape[0]
        return ref_shape[0] * _align4(k)
    return n


def to_native(kind, t):
    if kind == 'dw':                       # [C,1,3,3] -> [9][C]
        return t.reshape(t.shape[0], 9).t().contiguous()
    if kind == 'ohwi':                     # [O,I,kh,kw] -> [O][kh][kw][I]
        return t.permute(0, 2, 3, 1).contiguous()
    if kind == 'ohwi_pad4':                # same, each row zero-padded to a multiple of 4 (SPN conv1: 363 -> 364)
        O = t.shape[0]
        r = t.permute(0, 2, 3, 1).reshape(O, -1)
        out = t.new_zeros(O, _align4(r.shape[1]))
        out[:, :r.shape[1]] = r
        return out
    if kind.startswith('fc_chw'):          # Linear over a flattened NCHW map -> columns in NHWC order (SPN fc6/fc9)
        c, h, w = (int(v) for v in kind.split(':')[1].split('x'))
        return t.reshape(t.shape[0], c, h, w).permute(0, 2, 3, 1).reshape(t.shape[0], -1).contiguous()
    return t.contiguous()


def from_native(kind, flat, ref_shape):
    if kind == 'dw':
        C = ref_shape[0]
        return flat.view(9, C).t().reshape(ref_shape).contiguous()
    if kind == 'ohwi':
        O, I, kh, kw = ref_shape
        return flat.view(O, kh, kw, I).permute(0, 3, 1, 2).contiguous()
    if kind == 'ohwi_pad4':
        O, I, kh, kw = ref_shape
        return flat.view(O, -1)[:, :I * kh * kw].reshape(O, kh, kw, I).permute(0, 3, 1, 2).contiguous()
    if kind.startswith('fc_chw'):
        c, h, w = (int(v) for v in kind.split(':')[1].split('x'))
        return flat.view(ref_shape[0], h, w, c).permute(0, 3, 1, 2).reshape(ref_shape).contiguous()
    return flat.view(ref_shape).clone()


class ParamStore:
    """weights: list of (key, kind, ref_shape); bns: list of (prefix, C)."""

    def __init__(self, weights, bns, device):
        self.device = device
        self.entries = OrderedDict()
        off = 0
        for key, kind, shp in weights:
            e = Entry(key, kind, shp)
            e.off = off
            off = _align8(off + e.numel)
            self.entries[key] = e
        self.bns = list(bns)
        self.bn_index = {p: i for i, (p, _) in enumerate(self.bns)}
        self.bn_off = []
        c = 0
        for _, C in self.bns:
            self.bn_off.append(c)
            c += _align8(C)
        self.totC = c
        self.gamma_off = off
        self.beta_off = off + self.totC
        self.n = off + 2 * self.totC
        self.params = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.grads = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.bufs = torch.zeros(2 * self.totC, dtype=torch.float32, device=device)
        self.bufs[self.totC:] = 1.0
        self.nbt = torch.zeros(max(1, len(self.bns)), dtype=torch.int64, device=device)
        self.params[self.gamma_off:self.gamma_off + self.totC] = 1.0
        self.params_lowp = None          # bf16 mirror of `params` (GEMM weights of the --use_fp16 path), see enable_lowp()

    def enable_lowp(self):
        """allocate / refresh the bf16 mirror; the fused AdamW keeps it current afterwards (b200sp_adamw_step p_lowp)."""
        if self.params_lowp is None:
            self.params_lowp = torch.empty(self.n, dtype=torch.bfloat16, device=self.device)
        self.params_lowp.copy_(self.params)
        return self.params_lowp

    def wl_ptr(self, key):
        return self.params_lowp.data_ptr() + 2 * self.entries[key].off

    # ---- pointers ---------------------------------------------------------------------------
    def p_ptr(self, off):
        return self.params.data_ptr() + 4 * off

    def g_ptr(self, off):
        return self.grads.data_ptr() + 4 * off

    def w_ptr(self, key):
        return self.p_ptr(self.entries[key].off)

    def wg_ptr(self, key):
        return self.g_ptr(self.entries[key].off)

    def view(self, key, grads=False):
        e = self.entries[key]
        return (self.grads if grads else self.params)[e.off:e.off + e.numel]

    def bn_slices(self, i):
        o, C = self.bn_off[i], self.bns[i][1]
        return (slice(self.gamma_off + o, self.gamma_off + o + C), slice(self.beta_off + o, self.beta_off + o + C),
                slice(o, o + C), slice(self.totC + o, self.totC + o + C))

    # ---- reference-compatible state dict ------------------------------------------------------
    def ref_keys(self):
        """key -> ('w', entry) | ('bn', index, field)"""
        out = OrderedDict()
        for k, e in self.entries.items():
            out[k] = ('w', e)
        for i, (p, _) in enumerate(self.bns):
            for f in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked'):
                out[p + '.' + f] = ('bn', i, f)
        return out

    def state_dict(self, order=None):
        rk = self.ref_keys()
        keys = order if order is not None else list(rk)
        sd = OrderedDict()
        for k in keys:
            r = rk[k]
            if r[0] == 'w':
                e = r[1]
                sd[k] = from_native(e.kind, self.params[e.off:e.off + e.numel], e.ref_shape)
            else:
                g, b, rm, rv = self.bn_slices(r[1])
                f = r[2]
                if f == 'weight':
                    sd[k] = self.params[g].clone()
                elif f == 'bias':
                    sd[k] = self.params[b].clone()
                elif f == 'running_mean':
                    sd[k] = self.bufs[rm].clone()
                elif f == 'running_var':
                    sd[k] = self.bufs[rv].clone()
                else:
                    sd[k] = self.nbt[r[1]].clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        rk = self.ref_keys()
        missing = [k for k in rk if k not in sd]
        unexpected = [k for k in sd if k not in rk]
        if strict and (missing or unexpected):
            raise RuntimeError('Error(s) in loading state_dict: missing keys %s, unexpected keys %s'
                               % (missing[:8], unexpected[:8]))
        with torch.no_grad():
            for k, v in sd.items():
                if k not in rk:
                    continue
                r = rk[k]
                v = v.detach()
                if r[0] == 'w':
                    e = r[1]
                    if tuple(v.shape) != e.ref_shape:
                        raise RuntimeError('size mismatch for %s: %s vs %s' % (k, tuple(v.shape), e.ref_shape))
                    nat = to_native(e.kind, v.to(torch.float32)).reshape(-1)
                    self.params[e.off:e.off + e.numel].copy_(nat)
                else:
                    g, b, rm, rv = self.bn_slices(r[1])
                    f = r[2]
                    if f == 'weight':
                        self.params[g].copy_(v)
                    elif f == 'bias':
                        self.params[b].copy_(v)
                    elif f == 'running_mean':
                        self.bufs[rm].copy_(v)
                    elif f == 'running_var':
                        self.bufs[rv].copy_(v)
                    else:
                        self.nbt[r[1]] = int(v)
            if self.params_lowp is not None:
                self.params_lowp.copy_(self.params)
        return missing, unexpected

    def grad_dict(self):
        """reference-layout gradients (for parity tests)."""
        rk = self.ref_keys()
        out = OrderedDict()
        for k, r in rk.items():
            if r[0] == 'w':
                e = r[1]
                out[k] = from_native(e.kind, self.grads[e.off:e.off + e.numel], e.ref_shape)
            elif r[2] == 'weight':
                out[k] = self.grads[self.bn_slices(r[1])[0]].clone()
            elif r[2] == 'bias':
                out[k] = self.grads[self.bn_slices(r[1])[1]].clone()
        return out
