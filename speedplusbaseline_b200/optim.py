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

    def zero_grad(self, set_to_none=True):
        self.store.grads.zero_()

    def _device_hp(self):
        return L.AdamWHp.from_buffer_copy(self._hp.cpu().numpy().tobytes())

    def last_grad_norm(self):
        return self._device_hp().last_norm

    def device_step(self):
        """Number of optimizer steps taken, read from the device-resident counter: CUDA-graph replays advance it without
        Python's step() running, so the host copy is only the value last pushed."""
        self._step_host = int(self._device_hp().step)
        return self._step_host

    # ---- checkpoint format: torch.optim's own (reference utils.py:109-135 stores optimizer.state_dict()) ------------
    # {'state': {i: {'step', <per-parameter moment tensors in the REFERENCE shape>}}, 'param_groups': [{..., 'params': [0..n-1]}]}
    # where i enumerates the reference module's parameters() (= its state_dict order without buffers), so a
    # checkpoint.pth.tar written by the reference (or by this repo's --no_cuda path) resumes here and vice versa.
    _STATE_NAMES = {L.OPT_ADAMW: ('exp_avg', 'exp_avg_sq'), L.OPT_ADAM: ('exp_avg', 'exp_avg_sq'),
                    L.OPT_SGD: ('momentum_buffer', None), L.OPT_RMSPROP: ('square_avg', None)}

    def _ref_params(self):
        """[(flat slice, kind, reference shape)] in the reference's parameters() order."""
        st = self.store
        rk = st.ref_keys()
        order = getattr(st, 'key_order', None) or list(rk)
        out = []
        for k in order:
            r = rk[k]
            if r[0] == 'w':
                e = r[1]
                out.append((slice(e.off, e.off + e.numel), e.kind, e.ref_shape))
            elif r[2] in ('weight', 'bias'):
                g, b, _, _ = st.bn_slices(r[1])
                out.append((g if r[2] == 'weight' else b, 'plain', (st.bns[r[1]][1],)))
        return out

    def state_dict(self):
        from .params import from_native
        step = self.device_step()
        n1, n2 = self._STATE_NAMES[self.KIND]
        ref = self._ref_params()
        state = {}
        has_state = step > 0 and not (self.KIND == L.OPT_SGD and self.param_groups[0].get('momentum', 0) == 0)
        if has_state:
            for i, (sl, kind, shp) in enumerate(ref):
                d = {}
                if self.KIND != L.OPT_SGD:
                    d['step'] = torch.tensor(float(step))
                d[n1] = from_native(kind, self.exp_avg[sl], shp)
                if n2 is not None:
                    d[n2] = from_native(kind, self.exp_avg_sq[sl], shp)
                state[i] = d
        groups = []
        for g in self.param_groups[:1]:
            gg = {k: v for k, v in g.items() if k != 'params'}
            gg['params'] = list(range(len(ref)))
            groups.append(gg)
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        from .params import to_native
        s = sd['state']
        if sd.get('layout') == 'b200sp-flat':              # round-1 private layout: still readable
            self.exp_avg.copy_(s['exp_avg'])
            if self.exp_avg_sq is not None:
                self.exp_avg_sq.copy_(s['exp_avg_sq'])
            self._step_host = int(s['step'])
        else:
            n1, n2 = self._STATE_NAMES[self.KIND]
            ref = self._ref_params()
            if len(s) not in (0, len(ref)):
                raise ValueError('optimizer state has %d parameters, the model has %d' % (len(s), len(ref)))
            self.exp_avg.zero_()
            if self.exp_avg_sq is not None:
                self.exp_avg_sq.zero_()
            step = 0
            for i, (sl, kind, shp) in enumerate(ref):
                d = s.get(i, s.get(str(i)))
                if d is None:
                    continue
                if tuple(d[n1].shape) != tuple(shp):
                    raise ValueError('optimizer state %d: shape %s, expected %s' % (i, tuple(d[n1].shape), tuple(shp)))
                self.exp_avg[sl].copy_(to_native(kind, d[n1].to(torch.float32)).reshape(-1))
                if n2 is not None:
                    self.exp_avg_sq[sl].copy_(to_native(kind, d[n2].to(torch.float32)).reshape(-1))
                if 'step' in d:
                    step = max(step, int(float(d['step'])))
            self._step_host = step
        for g, gs in zip(self.param_groups, sd['param_groups']):
            g.update({k: v for k, v in gs.items() if k != 'params'})
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
