"""Fused clip + AdamW over the flat parameter buffer (b200sp_grad_sqnorm / b200sp_adamw_step).

Drop-in for the object `get_optimizer` returns in the reference (src/nets/build.py:60-78):
a torch.optim.Optimizer subclass, so StepLR (train.py:107), `.param_groups[i]['lr']`
(trainer.py:51-52), `.zero_grad(set_to_none=True)` and state_dict()/load_state_dict() work.
"""
import ctypes as C

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, store, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 clip_mode=0, max_norm=1.0, clip_value=1.0):
        super().__init__(list(params), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.store = store
        dev = store.params.device
        self.exp_avg = torch.zeros_like(store.params)
        self.exp_avg_sq = torch.zeros_like(store.params)
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
        key = (g['lr'], g['betas'], g['eps'], g['weight_decay'], self.clip_mode, self.max_norm, self.clip_value, self.grad_scale)
        if not force and key == self._last_pushed:
            return
        if not force:
            # keep the device-side step counter: read-modify-write only the hyper-parameter fields
            cur = self._hp.cpu().numpy().tobytes()
            C.memmove(C.addressof(self._hp_host), cur, C.sizeof(L.AdamWHp))
        h = self._hp_host
        h.lr, (h.beta1, h.beta2), h.eps, h.weight_decay = g['lr'], g['betas'], g['eps'], g['weight_decay']
        h.max_norm, h.clip_value, h.clip_mode, h.grad_scale = self.max_norm, self.clip_value, self.clip_mode, self.grad_scale
        if force:
            h.step, h.sqnorm, h.last_norm = self._step_host, 0.0, 0.0
        buf = (C.c_uint8 * C.sizeof(L.AdamWHp)).from_buffer_copy(bytes(h))
        self._staging.copy_(torch.frombuffer(buf, dtype=torch.uint8))
        self._hp.copy_(self._staging, non_blocking=False)
        self._last_pushed = key

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
        L.call('b200sp_adamw_step', st.params.data_ptr(), st.grads.data_ptr(), self.exp_avg.data_ptr(),
               self.exp_avg_sq.data_ptr(), st.params_lowp.data_ptr() if getattr(st, 'params_lowp', None) is not None else None, st.n, hp, sp)
        self._step_host += 1

    def zero_grad(self, set_to_none=True):
        self.store.grads.zero_()

    def last_grad_norm(self):
        cur = self._hp.cpu().numpy().tobytes()
        h = L.AdamWHp.from_buffer_copy(cur)
        return h.last_norm

    def state_dict(self):
        return {'state': {'step': self._step_host, 'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone()},
                'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups],
                'layout': 'b200sp-flat'}

    def load_state_dict(self, sd):
        s = sd['state']
        self.exp_avg.copy_(s['exp_avg'])
        self.exp_avg_sq.copy_(s['exp_avg_sq'])
        self._step_host = int(s['step'])
        for g, gs in zip(self.param_groups, sd['param_groups']):
            g.update(gs)
        self._push(force=True)
