"""RevGrad (DANN) -- same constructor / forward contract as /root/reference/src/nets/revgrad.py:58-96:
a KeypointRegressionNet under the `net.` prefix whose base[-1] output feeds, through a gradient
reversal layer (revgrad.py:36-56), the domain classifier conv1x1(320,1280)+ReLU -> AvgPool(7) ->
conv1x1(1280,1).  One KRNEngine (dann=True) owns all parameters in one flat store, so the fused
clip + AdamW covers both parts."""
import torch

from .. import _lib as L
from ..krn_engine import KRNEngine
from .park2019 import EngineModule, default_init


class _EnginePass(torch.autograd.Function):
    """Autograd holder for one RevGrad pass: outputs (loss | None, domain logits); backward receives
    both gradients at once and runs the CUDA backward of that pass."""

    @staticmethod
    def forward(ctx, anchor, fwd, bwd):
        ctx.bwd = bwd
        loss, z = fwd()
        ctx.has_loss = loss is not None
        if loss is None:
            loss = z.new_zeros(())
        return loss, z

    @staticmethod
    def backward(ctx, gl, gz):
        ctx.bwd(gl if ctx.has_loss else None, gz)
        return None, None, None


class RevGrad(EngineModule):
    def __init__(self, num_keypoints, device=None, dtype=L.F32, seed=None):
        super().__init__()
        self.nK = num_keypoints
        self.engine = KRNEngine(num_keypoints, prefix='net.', dann=True, device=device, dtype=dtype)
        self._register_store(self.engine.store, self.engine.key_order)
        default_init(self.engine.store, seed)
        if seed is None:
            from .park2019 import KeypointRegressionNet
            KeypointRegressionNet.load_imagenet_backbone(self)       # revgrad.py:64 builds the same pretrained KRN
        self._slot = 0

    def begin_step(self):
        """The reference calls the module twice per iteration (source, then target; dann.py:81,89) and
        backpropagates through both: each call of an iteration gets its own activation context."""
        self._slot = 0

    def forward(self, x, y=None, alpha=None):
        eng = self.engine
        x = x.contiguous().float()
        slot = self._slot
        self._slot += 1
        if y is None and alpha is None:
            cx = eng.forward(x, None, train=self.training, slot=slot)
            return cx.logits[:, 0::2].cpu(), cx.logits[:, 1::2].cpu()
        if y is not None:
            y = y.contiguous().float()
        state = {}
        neg_alpha = torch.tensor([-(float(alpha) if alpha is not None else 0.0)], dtype=torch.float32, device=x.device)

        def fwd():
            cx = state['cx'] = eng.forward(x, y, train=self.training, slot=slot)
            z = None
            if alpha is not None:
                z = eng.domain_forward(cx, 1.0).clone()      # label unused here: the caller computes the BCE
            loss = cx.loss3[0].clone() if y is not None else None
            if z is None:
                z = cx.loss3.new_zeros(cx.B)
            return loss, z

        def bwd(gl, gz):
            cx = state['cx']
            self.rebind_grads()
            fg = None
            if alpha is not None and gz is not None:
                cx.dom_dz.copy_(gz.reshape(-1))
                fg = eng.domain_backward(cx, neg_alpha)
            if gl is not None:
                if not (gl.numel() == 1 and float(gl) == 1.0):
                    cx.dlogits.mul_(gl)
                eng.backward(cx, feature_grad=fg, pose=True)
            elif fg is not None:
                eng.backward(cx, feature_grad=fg, pose=False)

        loss, z = _EnginePass.apply(self._plist[0], fwd, bwd)
        cx = state['cx']
        if y is not None:
            l3 = cx.loss3
            out1 = (loss, {'loss_x': float(l3[1]), 'loss_y': float(l3[2])})
        else:
            out1 = (cx.logits[:, 0::2].cpu(), cx.logits[:, 1::2].cpu())
        if alpha is None:
            return out1
        return out1, z.squeeze()
