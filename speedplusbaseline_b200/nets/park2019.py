"""KeypointRegressionNet -- same constructor / forward contract as
/root/reference/src/nets/park2019.py:100-165, executed by libb200sp kernels (krn_engine.KRNEngine)."""
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _lib as L
from ..krn_engine import KRNEngine


class _EngineLoss(torch.autograd.Function):
    """Autograd holder: forward runs the CUDA forward (+loss), backward runs the CUDA backward,
    which accumulates straight into the flat gradient buffer the parameters' .grad alias."""

    @staticmethod
    def forward(ctx, anchor, module, fwd, bwd):
        ctx.bwd = bwd
        return fwd()

    @staticmethod
    def backward(ctx, gout):
        ctx.bwd(gout)
        return None, None, None, None


class EngineModule(nn.Module):
    """nn.Module facade over an engine with a flat ParamStore: exposes per-key Parameters that
    alias the flat buffers (so model.parameters(), clip_grad_norm_, GradScaler.unscale_ work) and
    a reference-compatible state_dict."""

    def _register_store(self, store, key_order):
        self._store, self._key_order = store, key_order
        store.key_order = list(key_order)       # reference parameters() order for the optimizer checkpoint (optim.py)
        self._plist = nn.ParameterList()
        for k, e in store.entries.items():
            p = nn.Parameter(store.params[e.off:e.off + e.numel])
            p.grad = store.grads[e.off:e.off + e.numel]
            self._plist.append(p)
        for off in (store.gamma_off, store.beta_off):
            p = nn.Parameter(store.params[off:off + store.totC])
            p.grad = store.grads[off:off + store.totC]
            self._plist.append(p)

    def rebind_grads(self):
        """re-alias .grad to the flat gradient buffer (after zero_grad(set_to_none=True))."""
        st, i = self._store, 0
        for k, e in st.entries.items():
            self._plist[i].grad = st.grads[e.off:e.off + e.numel]
            i += 1
        for off in (st.gamma_off, st.beta_off):
            self._plist[i].grad = st.grads[off:off + st.totC]
            i += 1

    def state_dict(self, *args, **kwargs):
        return self._store.state_dict(self._key_order)

    def load_state_dict(self, state_dict, strict=True):
        self._store.load_state_dict(state_dict, strict)

    def to(self, *args, **kwargs):          # parameters already live on the engine's device
        return self

    def grad_dict(self):
        return self._store.grad_dict()


def default_init(store, seed=None, kaiming_prefixes=('base.', 'net.base.')):
    """Random init with the reference's distributions: torchvision MobileNetV2 body = kaiming-normal
    fan_out convs (mobilenetv2.py:143-153); everything else = nn.Conv2d/nn.Linear defaults
    (kaiming-uniform a=sqrt(5) weights, U(+-1/sqrt(fan_in)) biases); BN = (1, 0).  The reference loads
    ImageNet weights into the body (park2019.py:107), which cannot be downloaded offline; use
    load_state_dict() with a torchvision checkpoint for that."""
    import math
    g = torch.Generator(device='cpu')
    if seed is not None:
        g.manual_seed(seed)
    else:
        g.seed()
    sd, fan_ins = {}, {}
    for k, e in store.entries.items():
        shp = e.ref_shape
        if len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            fan_ins[k.rsplit('.', 1)[0]] = fan_in
            if k.startswith(kaiming_prefixes) and len(shp) == 4:
                fan_out = shp[0] * shp[2] * shp[3]
                sd[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_out)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    for k, e in store.entries.items():
        if len(e.ref_shape) == 1:
            bound = 1.0 / math.sqrt(fan_ins.get(k.rsplit('.', 1)[0], 1))
            sd[k] = (torch.rand(e.ref_shape, generator=g) * 2 - 1) * bound
    store.load_state_dict(sd, strict=False)


class KeypointRegressionNet(EngineModule):
    def __init__(self, num_keypoints, device=None, dtype=L.F32, seed=None, _prefix='', _dann=False, tf32_gemm=False):
        super().__init__()
        self.nK = num_keypoints
        self.engine = KRNEngine(num_keypoints, prefix=_prefix, dann=_dann, device=device, dtype=dtype, tf32_gemm=tf32_gemm)
        self._register_store(self.engine.store, self.engine.key_order)
        default_init(self.engine.store, seed)
        if seed is None:
            self.load_imagenet_backbone()

    def load_imagenet_backbone(self, checkpoint=None):
        """park2019.py:107 `models.mobilenet_v2(pretrained=True)`: the ImageNet backbone.  There is no network here, so the
        weights are taken from where torchvision would have cached its download
        ($TORCH_HOME/hub/checkpoints/mobilenet_v2-b0353104.pth) or from an explicit path / state dict; when neither exists
        the backbone keeps its random initialisation (warned).  Seeded constructions (tests, bench) never load."""
        import logging
        import os
        from ..importers import load_mobilenetv2_backbone
        if checkpoint is None:
            hub = os.path.join(os.environ.get('TORCH_HOME', os.path.join(os.path.expanduser('~'), '.cache', 'torch')), 'hub', 'checkpoints')
            for fn in ('mobilenet_v2-b0353104.pth', 'mobilenet_v2-7ebf99e0.pth'):
                if os.path.exists(os.path.join(hub, fn)):
                    checkpoint = os.path.join(hub, fn)
                    break
        if checkpoint is None:
            logging.getLogger(__name__).warning('   - no cached torchvision MobileNetV2 checkpoint: backbone keeps its random initialisation')
            return []
        return load_mobilenetv2_backbone(self, checkpoint)

    def forward(self, x, y=None):
        eng = self.engine
        x = x.contiguous().float()
        if y is not None:
            # TRAINING (park2019.py:146-162)
            y = y.contiguous().float()
            state = {}

            def fwd():
                state['cx'] = eng.forward(x, y, train=self.training)
                return state['cx'].loss3[0].clone()

            def bwd(gout):
                cx = state['cx']
                if not (gout.numel() == 1 and float(gout) == 1.0):
                    cx.dlogits.mul_(gout)
                self.rebind_grads()
                eng.backward(cx)

            loss = _EngineLoss.apply(self._plist[0], self, fwd, bwd)
            l3 = state['cx'].loss3
            sm = {'loss_x': float(l3[1]), 'loss_y': float(l3[2])}
            return loss, sm
        cx = eng.forward(x, None, train=self.training)
        logits = cx.logits
        return logits[:, 0::2].cpu(), logits[:, 1::2].cpu()
