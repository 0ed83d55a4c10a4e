"""DANN epoch loop with the reference's signature (/root/reference/src/core/dann.py:38-117) on a
fused step: per iteration TWO KRN forwards (source with labels, target without; dann.py:81,89), the
domain classifier + BCE on both, ONE backward through both graphs with the gradient reversal, global
norm clip and AdamW -- all libb200sp launches, captured in CUDA graphs.  alpha (dann.py:77-78) lives in
a device scalar so the captured graph is reused while it changes every iteration."""
import time

import numpy as np
import torch

from .. import _lib as L
from ..dist import GradSync, world_size
from ..utils import AverageMeter, report_progress


def dann_alpha(idx, epoch, n_batches, max_epochs):
    """dann.py:77-78"""
    p = float(idx + epoch * n_batches) / max_epochs / n_batches
    return 2. / (1. + np.exp(-10 * p)) - 1


class DANNTrainStep:
    """step(source, label, target, alpha) -> device tensor losses = (pose, domain_source, domain_target).
    All inputs are device tensors: source/target [B,3,224,224] fp32, label [B,2,K] fp32."""

    def __init__(self, model, optimizer, use_graph=True, world_size=1, process_group=None):
        self.model, self.opt, self.use_graph = model, optimizer, use_graph
        self.world, self.pg = world_size, process_group
        self.sync = GradSync(world_size, process_group)
        if world_size > 1:
            optimizer.grad_scale = self.sync.grad_scale      # 1/world folded into the AdamW kernel
        dev = model.engine.device
        self.neg_alpha = torch.zeros(1, dtype=torch.float32, device=dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self._graphs = self._static = self._sig = None

    def _fwd_bwd(self, source, label, target):
        eng = self.model.engine
        eng.store.grads.zero_()
        cs = eng.forward(source, label, train=True, slot=0)
        eng.domain_forward(cs, 1.0)                                  # dann.py:85-87 (1: source)
        ct = eng.forward(target, None, train=True, slot=1)           # dann.py:89 (updates BN stats of extras too)
        eng.domain_forward(ct, 0.0)                                  # dann.py:90-92 (0: target)
        # loss = pose + domain_source + domain_target (dann.py:95): unit weights, one backward
        fs = eng.domain_backward(cs, self.neg_alpha)
        eng.backward(cs, feature_grad=fs, pose=True)
        ft = eng.domain_backward(ct, self.neg_alpha)
        eng.backward(ct, feature_grad=ft, pose=False)
        self.losses[0:1].copy_(cs.loss3[0:1])
        self.losses[1:2].copy_(cs.dom_loss)
        self.losses[2:3].copy_(ct.dom_loss)
        return self.losses

    def _allreduce(self):
        self.sync.allreduce(self.model.engine.store.grads)

    def _set_alpha(self, alpha):
        # the value travels as a kernel argument of the fill: no reused pinned staging scalar that a later iteration
        # could overwrite before an earlier asynchronous copy has read it
        self.neg_alpha.fill_(-float(alpha))

    def eager(self, source, label, target, alpha):
        self.opt.sync_hyperparams()
        self._set_alpha(alpha)
        out = self._fwd_bwd(source, label, target)
        self._allreduce()
        self.opt.step(sync=False)
        return out

    def _capture(self, source, label, target):
        self._static = tuple(torch.empty_like(t) for t in (source, label, target))
        for s, t in zip(self._static, (source, label, target)):
            s.copy_(t)
        st = self.model.engine.store
        snap = (st.bufs.clone(), st.nbt.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                   # warm-up: allocates every activation buffer
            self._fwd_bwd(*self._static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._fwd_bwd(*self._static)
        with torch.cuda.graph(g2):
            self.opt.step(sync=False)
        st.bufs.copy_(snap[0]); st.nbt.copy_(snap[1])      # warm-up/capture must not disturb BN buffers
        self._graphs = (g1, g2)
        self._sig = tuple(tuple(t.shape) for t in (source, label, target))

    def step(self, source, label, target, alpha):
        if not self.use_graph:
            return self.eager(source, label, target, alpha)
        sig = tuple(tuple(t.shape) for t in (source, label, target))
        if self._graphs is None or sig != self._sig:
            self._capture(source, label, target)
        self.opt.sync_hyperparams()
        self._set_alpha(alpha)
        for s, t in zip(self._static, (source, label, target)):
            s.copy_(t, non_blocking=True)
        self._graphs[0].replay()
        self._allreduce()
        self._graphs[1].replay()
        return self.losses


def train_dann_single_epoch_krn(epoch, cfg, model, dataloader_source, dataloader_target,
                                optimizer, writer, device, scaler=None):
    """Same signature and behaviour as reference dann.py:38-117 (DANN is fp32-only there, adapt.py:99-101)."""
    training_time_meter = AverageMeter('ms')
    loss_pose_meter = AverageMeter('-')
    loss_source_meter = AverageMeter('-')
    loss_target_meter = AverageMeter('-')
    model.train()
    for pg in optimizer.param_groups:
        lr = pg['lr']
    batches = zip(dataloader_source, dataloader_target)
    n_batches = min(len(dataloader_source), len(dataloader_target))
    stepper = getattr(model, '_dann_step', None)
    if stepper is None or stepper.opt is not optimizer:
        stepper = DANNTrainStep(model, optimizer, use_graph=getattr(cfg, 'use_graph', True), world_size=world_size())
        model._dann_step = stepper
    pending = None
    for idx, ((source, label), target) in enumerate(batches):
        B = source.size(0)
        ts = time.time()
        source = source.to(device, non_blocking=True).float().contiguous()
        label = label.to(device, non_blocking=True).float().contiguous()
        target = target.to(device, non_blocking=True).float().contiguous()
        alpha = dann_alpha(idx, epoch, n_batches, cfg.max_epochs)
        losses = stepper.step(source, label, target, alpha)
        if pending is not None:                         # previous iteration's losses: wait for THAT copy only
            hl, pb, ev = pending
            ev.synchronize()
            loss_pose_meter.update(float(hl[0]), pb)
            loss_source_meter.update(float(hl[1]), pb)
            loss_target_meter.update(float(hl[2]), pb)
        host = torch.empty(3, pin_memory=True)
        host.copy_(losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending = (host, B, ev)
        training_time_meter.update((time.time() - ts) * 1000, B)
        report_progress(epoch=epoch, lr=lr, epoch_iter=idx + 1, epoch_size=n_batches,
                        time=training_time_meter, is_train=True,
                        loss_pose=loss_pose_meter, loss_source=loss_source_meter, loss_target=loss_target_meter)
    if pending is not None:
        torch.cuda.synchronize()
        hl, pb, _ = pending
        loss_pose_meter.update(float(hl[0]), pb)
        loss_source_meter.update(float(hl[1]), pb)
        loss_target_meter.update(float(hl[2]), pb)
    if writer is not None:
        writer.add_scalar('train/loss_pose', loss_pose_meter.avg, epoch)
        writer.add_scalar('train/loss_source', loss_source_meter.avg, epoch)
        writer.add_scalar('train/loss_target', loss_target_meter.avg, epoch)
