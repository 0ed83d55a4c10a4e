"""Fused clip + optimizer step over the flat parameter buffer (b200sp_grad_sqnorm / b200sp_adamw_step /
b200sp_optim_step).

Drop-in for the object `get_optimizer` returns in the reference (src/nets/build.py:60-78):
a torch.optim.Optimizer subclass, so StepLR (train.py:107), `.param_groups[i]['lr']`
(trainer.py:51-52), `.zero_grad(set_to_none=True)` and state_dict()/load_state_dict() work.
FusedAdamW is the north-star path (`--optimizer adamw`); FusedSGD / FusedRMSprop / FusedAdam cover the other three
choices of build.py:63-71 with the same device-resident hyper-parameter block (CUDA-graph replayable).
"""
import ctypes as C

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    KIND = L.OPT_ADAMW

    def __init__(self, store, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 clip_mode=0, max_norm=1.0, clip_value=1.0):
        self._init_flat(store, params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay),
                        clip_mode, max_norm, clip_value)

    def _init_flat(self, store, params, defaults, clip_mode, max_norm, clip_value):
        torch.optim.Optimizer.__init__(self, list(params), defaults)
        self.store = store
        dev = store.params.device
        self.exp_avg = torch.zeros_like(store.params)          # s1: exp_avg | momentum_buffer | square_avg
        self.exp_avg_sq = torch.zeros_like(store.params) if self.KIND in (L.OPT_ADAMW, L.OPT_ADAM) else None
        self.clip_mode, self.max_norm, self.clip_value = clip_mode, max_norm, clip_value
        self.grad_scale = 1.0
        self._hp_host = L.AdamWHp()
        self._hp = torch.zeros(C.sizeof(L.AdamWHp), dtype=torch.uint8, device=dev)
        self._staging = torch.zeros(C.sizeof(L.AdamWHp), dtype=torch.uint8).pin_memory()
        self._step_host = 0
        self._last_pushed = None
        self._push(force=True)

    def _push(self, force=False):
        g = self.param_groups[0]
        g_betas, g_eps = self._betas_eps(g)
        key = (g['lr'], g_betas, g_eps, g['weight_decay'], self.clip_mode, self.max_norm, self.clip_value, self.grad_scale)
        if not force and key == self._last_pushed:
            return
        if not force:
            # keep the device-side step counter: read-modify-write only the hyper-parameter fields
            cur = self._hp.cpu().numpy().tobytes()
            C.memmove(C.addressof(self._hp_host), cur, C.sizeof(L.AdamWHp))
        h = self._hp_host
        h.lr, (h.beta1, h.beta2), h.eps, h.weight_decay = g['lr'], g_betas, g_eps, g['weight_decay']
        h.max_norm, h.clip_value, h.clip_mode, h.grad_scale = self.max_norm, self.clip_value, self.clip_mode, self.grad_scale
        if force:
            h.step, h.sqnorm, h.last_norm = self._step_host, 0.0, 0.0
        buf = (C.c_uint8 * C.sizeof(L.AdamWHp)).from_buffer_copy(bytes(h))
        self._staging.copy_(torch.frombuffer(buf, dtype=torch.uint8))
        self._hp.copy_(self._staging, non_blocking=False)
        self._last_pushed = key

    @staticmethod
    def _betas_eps(g):
        """(beta1, beta2), eps of the device block from a param_group (overridden per optimizer kind)."""
        return tuple(g['betas']), g['eps']

    def sync_hyperparams(self):
        """Call outside CUDA-graph capture after lr (StepLR) changes."""
        self._push()

    @torch.no_grad()
    def step(self, closure=None, sync=True):
        if sync:
            self._push()
        st, sp = self.store, L.stream_ptr()
        hp = self._hp.data_ptr()
        if self.clip_mode == 1:
            L.call('b200sp_grad_sqnorm', st.grads.data_ptr(), st.n, hp, sp)
        lowp = st.params_lowp.data_ptr() if getattr(st, 'params_lowp', None) is not None else None
        if self.KIND == L.OPT_ADAMW:
            L.call('b200sp_adamw_step', st.params.data_ptr(), st.grads.data_ptr(), self.exp_avg.data_ptr(),
                   self.exp_avg_sq.data_ptr(), lowp, st.n, hp, sp)
        else:
            L.call('b200sp_optim_step', self.KIND, st.params.data_ptr(), st.grads.data_ptr(), self.exp_avg.data_ptr(),
                   L.ptr(self.exp_avg_sq), lowp, st.n, hp, sp)
        self._step_host += 1

    def zero_grad(self, set_to_none=True):
        self.store.grads.zero_()

    def last_grad_norm(self):
        cur = self._hp.cpu().numpy().tobytes()
        h = L.AdamWHp.from_buffer_copy(cur)
        return h.last_norm

    def state_dict(self):
        return {'state': {'step': self._step_host, 'exp_avg': self.exp_avg.clone(),
                          'exp_avg_sq': None if self.exp_avg_sq is None else self.exp_avg_sq.clone()},
                'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups],
                'layout': 'b200sp-flat'}

    def load_state_dict(self, sd):
        s = sd['state']
        self.exp_avg.copy_(s['exp_avg'])
        if self.exp_avg_sq is not None:
            self.exp_avg_sq.copy_(s['exp_avg_sq'])
        self._step_host = int(s['step'])
        for g, gs in zip(self.param_groups, sd['param_groups']):
            g.update(gs)
        self._push(force=True)


class FusedSGD(FusedAdamW):
    """torch.optim.SGD(param, lr, momentum, weight_decay) of build.py:63-65 (dampening 0, no nesterov)."""
    KIND = L.OPT_SGD

    def __init__(self, store, params, lr=1e-3, momentum=0.0, weight_decay=0.0, clip_mode=0, max_norm=1.0, clip_value=1.0):
        self._init_flat(store, params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay), clip_mode, max_norm, clip_value)

    @staticmethod
    def _betas_eps(g):
        return (g['momentum'], 0.0), 0.0


class FusedRMSprop(FusedAdamW):
    """torch.optim.RMSprop(param, lr, alpha, weight_decay) of build.py:66-68 (eps 1e-8, momentum 0, not centered)."""
    KIND = L.OPT_RMSPROP

    def __init__(self, store, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0.0, clip_mode=0, max_norm=1.0, clip_value=1.0):
        self._init_flat(store, params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay), clip_mode, max_norm, clip_value)

    @staticmethod
    def _betas_eps(g):
        return (g['alpha'], 0.0), g['eps']


class FusedAdam(FusedAdamW):
    """torch.optim.Adam(param, lr, betas, weight_decay) of build.py:69-71 (coupled L2 decay, eps 1e-8, no amsgrad)."""
    KIND = L.OPT_ADAM

    def __init__(self, store, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_mode=0, max_norm=1.0,
                 clip_value=1.0):
        self._init_flat(store, params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay), clip_mode, max_norm, clip_value)
