"""helpers for kernel-level GPU tests: build b200sp structs straight from torch tensors."""
import ctypes as C

import torch

from speedplusbaseline_b200 import _lib as L


def sp():
    return L.stream_ptr()


def vt_plain(t):
    return L.VTensor(t.data_ptr(), None, None, None, None, L.VT_PLAIN, 0)


def vt_bnact(y, scale, shift, act):
    return L.VTensor(y.data_ptr(), None, scale.data_ptr(), shift.data_ptr(), None, L.VT_BNACT, act)


def vt_dy(g, y, cA, cB, cC):
    return L.VTensor(g.data_ptr(), y.data_ptr(), cA.data_ptr(), cB.data_ptr(), cC.data_ptr(), L.VT_DY, 0)


class BnF:
    """forward-BN epilogue workspace for C channels"""

    def __init__(self, C_, gamma=None, beta=None, dev='cuda'):
        f = dict(device=dev, dtype=torch.float32)
        self.sum = torch.zeros(C_, device=dev, dtype=torch.float64)
        self.sumsq = torch.zeros(C_, device=dev, dtype=torch.float64)
        self.ticket = torch.zeros(4, device=dev, dtype=torch.int32)
        self.gamma = gamma if gamma is not None else torch.ones(C_, **f)
        self.beta = beta if beta is not None else torch.zeros(C_, **f)
        self.rm, self.rv = torch.zeros(C_, **f), torch.ones(C_, **f)
        self.scale, self.shift, self.mean, self.rstd = (torch.zeros(C_, **f) for _ in range(4))
        self.s = L.BnFwd(self.sum.data_ptr(), self.sumsq.data_ptr(), self.ticket.data_ptr(), self.gamma.data_ptr(),
                         self.beta.data_ptr(), self.rm.data_ptr(), self.rv.data_ptr(), self.scale.data_ptr(),
                         self.shift.data_ptr(), self.mean.data_ptr(), self.rstd.data_ptr(), 0.1, 1e-5)

    def ref(self):
        return C.byref(self.s)


class BnB:
    def __init__(self, y, scale, shift, mean, rstd, act, stats=True):
        dev, C_ = y.device, scale.numel()
        f = dict(device=dev, dtype=torch.float32)
        self.s1 = torch.zeros(C_, device=dev, dtype=torch.float64)
        self.s2 = torch.zeros(C_, device=dev, dtype=torch.float64)
        self.ticket = torch.zeros(4, device=dev, dtype=torch.int32)
        self.cA, self.cB, self.cC, self.dgamma, self.dbeta = (torch.zeros(C_, **f) for _ in range(5))
        self.keep = (y, scale, shift, mean, rstd)
        self.s = L.BnBwd(self.s1.data_ptr() if stats else None, self.s2.data_ptr() if stats else None,
                         self.ticket.data_ptr(), y.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                         rstd.data_ptr(), self.cA.data_ptr(), self.cB.data_ptr(), self.cC.data_ptr(),
                         self.dgamma.data_ptr(), self.dbeta.data_ptr(), act, 0)

    def ref(self):
        return C.byref(self.s)


def act_t(z, act):
    if act == L.ACT_RELU:
        return torch.relu(z)
    if act == L.ACT_RELU6:
        return torch.clamp(z, 0, 6)
    if act == L.ACT_LEAKY02:
        return torch.nn.functional.leaky_relu(z, 0.2)
    return z


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
