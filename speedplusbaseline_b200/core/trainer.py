"""Epoch loops with the reference's signatures (/root/reference/src/core/trainer.py:41-199) on top
of a fused, CUDA-graph-captured training step.

Reference step (trainer.py:64-98): H2D -> [style-aug] -> forward(+loss, 2 host syncs) -> zero_grad ->
backward -> clip_grad_norm_(1.0) -> AdamW.  Here the whole device part is ONE graph launch:
memset(grads) -> forward kernels -> backward kernels -> [NCCL allreduce] -> sqnorm -> AdamW;
losses are read back asynchronously, so there is no per-iteration host sync.
"""
import logging
import random
import time

import torch

from .. import _lib as L
from ..dist import GradSync, world_size
from ..utils import AverageMeter, report_progress

logger = logging.getLogger("Training")


class KRNTrainStep:
    """One fused KRN training iteration.  `step(images, target)` consumes device tensors
    ([B,3,H,W] fp32 in [0,1], [B,2,K] fp32) and returns the device tensor loss3 = (loss, loss_x, loss_y)."""

    def __init__(self, model, optimizer, use_graph=True, world_size=1, process_group=None):
        self.model, self.opt, self.use_graph = model, optimizer, use_graph
        self.world, self.pg = world_size, process_group
        self.sync = GradSync(world_size, process_group)
        if world_size > 1:
            optimizer.grad_scale = self.sync.grad_scale      # 1/world folded into the AdamW kernel
        self._graphs = None
        self._static = None
        self._sig = None

    # the un-captured sequence (also used for warm-up and as the eager fallback for odd batch sizes)
    def _fwd_bwd(self, images, target):
        eng = self.model.engine
        eng.store.grads.zero_()
        cx = eng.forward(images, target, train=True)
        eng.backward(cx)
        return cx

    def _update(self):
        self.opt.step(sync=False)

    def _allreduce(self):
        self.sync.allreduce(self.model.engine.store.grads)

    def eager(self, images, target):
        self.opt.sync_hyperparams()
        cx = self._fwd_bwd(images, target)
        self._allreduce()
        self._update()
        return cx.loss3

    def _capture(self, images, target):
        self._static = (torch.empty_like(images), torch.empty_like(target))
        self._static[0].copy_(images)
        self._static[1].copy_(target)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):          # warm-up allocates every activation buffer (step() snapshots / restores the BN buffers around this)
            cx = self._fwd_bwd(*self._static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        import os
        if self.world > 1 and os.environ.get('B200SP_GRAPH_NCCL', '0') == '1':
            # opt-in (B200SP_GRAPH_NCCL=1).  Measured at N=2 (profiles/r2_ddp_n2.txt): no gain over the eager all-reduce between two
            # graphs (6.31 vs 6.29 ms) and the process hangs in destroy_process_group afterwards, so it is NOT the default.
            # ONE graph for the whole step: the NCCL all-reduce of the flat gradient buffer is captured between the backward
            # kernels and the fused update (NCCL >= 2.9 supports stream capture), so a replay has no host-side gap around the
            # collective.  An eager all-reduce first: communicator set-up must not happen under capture.
            self._allreduce()
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                cx = self._fwd_bwd(*self._static)
                self._allreduce()
                self._update()
            self._graphs = (g1, None)
        else:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                cx = self._fwd_bwd(*self._static)
            with torch.cuda.graph(g2):
                self._update()
            self._graphs = (g1, g2)
        self._loss3 = cx.loss3
        self._sig = (tuple(images.shape), tuple(target.shape))

    def step(self, images, target):
        if not self.use_graph:
            return self.eager(images, target)
        sig = (tuple(images.shape), tuple(target.shape))
        if self._graphs is None or sig != self._sig:
            # capture needs a warm-up forward/backward that must not disturb model state
            st = self.model.engine.store
            snap = (st.bufs.clone(), st.nbt.clone())
            self._capture(images, target)
            st.bufs.copy_(snap[0]); st.nbt.copy_(snap[1])
        self.opt.sync_hyperparams()
        self._static[0].copy_(images, non_blocking=True)
        self._static[1].copy_(target, non_blocking=True)
        self._graphs[0].replay()
        if self._graphs[1] is not None:
            self._allreduce()
            self._graphs[1].replay()
        return self._loss3


def train_single_epoch_krn(epoch, cfg, model, data_loader, optimizer,
                           writer, device, styleAugmentor=None, scaler=None):
    """Same signature and behaviour as reference trainer.py:41-112 (scaler is accepted; bf16 needs no
    loss scaling so it is only used as the 'mixed precision on' flag)."""
    training_time_meter = AverageMeter('ms')
    loss_x_meter = AverageMeter('-')
    loss_y_meter = AverageMeter('-')
    model.train()
    for pg in optimizer.param_groups:
        lr = pg['lr']
    stepper = getattr(model, '_train_step', None)
    if stepper is None or stepper.opt is not optimizer:
        stepper = KRNTrainStep(model, optimizer, use_graph=getattr(cfg, 'use_graph', True), world_size=world_size())
        model._train_step = stepper
    pending = None
    for idx, (images, target) in enumerate(DevicePrefetcher(data_loader, device)):
        start = time.time()
        B = images.shape[0]
        if styleAugmentor is not None and random.random() < cfg.texture_ratio:
            images = styleAugmentor(images)
        loss3 = stepper.step(images, target)
        # read the PREVIOUS iteration's losses (already complete) instead of syncing on this one
        if pending is not None:
            l3, pb, pev = pending
            pev.synchronize()                           # waits for the PREVIOUS step's 12-byte copy only
            loss_x_meter.update(float(l3[1]), pb)
            loss_y_meter.update(float(l3[2]), pb)
        host = torch.empty(3, pin_memory=True)
        host.copy_(loss3, non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        pending = (host, B, ev)
        training_time_meter.update((time.time() - start) * 1000, B)
        report_progress(epoch=epoch, lr=lr, epoch_iter=idx + 1, epoch_size=len(data_loader),
                        time=training_time_meter, is_train=True, loss_x=loss_x_meter, loss_y=loss_y_meter)
    if pending is not None:
        torch.cuda.synchronize()
        loss_x_meter.update(float(pending[0][1]), pending[1])
        loss_y_meter.update(float(pending[0][2]), pending[1])
    if writer is not None:
        writer.add_scalar('train/loss_x', loss_x_meter.avg, epoch)
        writer.add_scalar('train/loss_y', loss_y_meter.avg, epoch)


class DevicePrefetcher:
    """Wraps a loader of pinned HOST batches: the host->device copy of batch i+1 runs on a copy stream while the
    GPU computes batch i (the reference copies synchronously in the loop, trainer.py:64-65).  Two device buffer sets;
    a buffer is refilled only after the step that consumed it has been enqueued (event), so nothing is overwritten
    early.  Yields tuples of device tensors valid until the next-but-one iteration."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device
        self.stream = torch.cuda.Stream(device=device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]

    def __len__(self):
        return len(self.loader)

    def _issue(self, slot, batch):
        batch = batch if isinstance(batch, (tuple, list)) else (batch,)
        if self.bufs[slot] is None or any(tuple(b.shape) != tuple(t.shape) for b, t in zip(self.bufs[slot], batch)):
            self.bufs[slot] = [torch.empty(t.shape, dtype=torch.float32, device=self.device) for t in batch]
        if self.consumed[slot] is not None:
            self.stream.wait_event(self.consumed[slot])
        if any(t.is_cuda for t in batch):
            # batches produced ON the device (datasets/raw.py:DeviceBatchLoader runs the transform stack on the current
            # stream): the copy stream must not read them before those kernels have finished
            self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for b, t in zip(self.bufs[slot], batch):
                b.copy_(t, non_blocking=True)
            self.ready[slot].record(self.stream)

    def mark_consumed(self, slot):
        ev = torch.cuda.Event()
        ev.record()
        self.consumed[slot] = ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = next(it)
        except StopIteration:
            return
        slot = 0
        self._issue(slot, nxt)
        while True:
            torch.cuda.current_stream().wait_event(self.ready[slot])
            cur, cur_slot = self.bufs[slot], slot
            try:
                nxt = next(it)
                self._issue(slot ^ 1, nxt)
                more = True
            except StopIteration:
                more = False
            yield tuple(cur)
            self.mark_consumed(cur_slot)        # the consumer has enqueued its work on the current stream
            if not more:
                return
            slot ^= 1


class SPNTrainStep:
    """One fused SPN training iteration (reference trainer.py:137-186): forward, CE(class) + 10 CE(weights),
    backward, clip_grad_value_(1.0) + AdamW (one launch, optim.FusedAdamW clip_mode=2).
    step(images [B,3,227,227], yClasses [B,N], yWeights [B,N]) -> device tensor (loss_class, loss_regress)."""

    def __init__(self, model, optimizer, use_graph=True, world_size=1, process_group=None):
        self.model, self.opt, self.use_graph = model, optimizer, use_graph
        self.world, self.pg = world_size, process_group
        self.sync = GradSync(world_size, process_group)
        if world_size > 1:
            optimizer.grad_scale = self.sync.grad_scale
        self._graphs = self._static = self._sig = None
        # dropout masks: counter-based generator keyed by (seed, rank) with a DEVICE-resident step counter, so the captured
        # graph draws a new mask per replay, ranks differ, and a resumed run continues the sequence (engine.drop_ctr)
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.model.engine.base_seed = (int(getattr(model, 'seed', 0) or 0) * 1000003 + rank) & 0x7fffffff

    def _fwd_bwd(self, images, yc, yw):
        eng = self.model.engine
        eng.store.grads.zero_()
        eng.forward(images, yc, yw, train=True)
        eng.backward()
        return eng.loss2

    def eager(self, images, yc, yw):
        self.opt.sync_hyperparams()
        out = self._fwd_bwd(images, yc, yw)
        self.sync.allreduce(self.model.engine.store.grads)
        self.opt.step(sync=False)
        return out

    def step(self, images, yc, yw):
        if not self.use_graph:
            return self.eager(images, yc, yw)
        sig = tuple(tuple(t.shape) for t in (images, yc, yw))
        if self._graphs is None or sig != self._sig:
            self._static = tuple(torch.empty_like(t) for t in (images, yc, yw))
            for s, t in zip(self._static, (images, yc, yw)):
                s.copy_(t)
            ctr = self.model.engine.drop_ctr.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._fwd_bwd(*self._static)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._fwd_bwd(*self._static)
            self.model.engine.drop_ctr.copy_(ctr)          # the warm-up pass is not a training step
            with torch.cuda.graph(g2):
                self.opt.step(sync=False)
            self._graphs, self._sig = (g1, g2), sig
        self.opt.sync_hyperparams()
        for s, t in zip(self._static, (images, yc, yw)):
            s.copy_(t, non_blocking=True)
        self._graphs[0].replay()
        self.sync.allreduce(self.model.engine.store.grads)
        self._graphs[1].replay()
        return self.model.engine.loss2


def train_single_epoch_spn(epoch, cfg, model, data_loader, optimizer,
                           writer, device, styleAugmentor=None, scaler=None):
    """Same signature and behaviour as reference trainer.py:114-199."""
    training_time_meter = AverageMeter('ms')
    loss_class_meter = AverageMeter('-')
    loss_weight_meter = AverageMeter('-')
    model.train()
    for pg in optimizer.param_groups:
        lr = pg['lr']
    stepper = getattr(model, '_train_step', None)
    if stepper is None or stepper.opt is not optimizer:
        stepper = SPNTrainStep(model, optimizer, use_graph=getattr(cfg, 'use_graph', True), world_size=world_size())
        model._train_step = stepper
    pending = None
    for idx, (images, yClasses, yWeights) in enumerate(data_loader):
        start = time.time()
        B = images.shape[0]
        images = images.to(device, non_blocking=True).float().contiguous()
        yClasses = yClasses.to(device, non_blocking=True).float().contiguous()
        yWeights = yWeights.to(device, non_blocking=True).float().contiguous()
        if styleAugmentor is not None and random.random() < cfg.texture_ratio:
            images = styleAugmentor(images)
        loss2 = stepper.step(images, yClasses, yWeights)
        if pending is not None:
            hl, pb, pev = pending
            pev.synchronize()
            loss_class_meter.update(float(hl[0]), pb)
            loss_weight_meter.update(float(hl[1]), pb)
        host = torch.empty(2, pin_memory=True)
        host.copy_(loss2, non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        pending = (host, B, ev)
        training_time_meter.update((time.time() - start) * 1000, B)
        report_progress(epoch=epoch, lr=lr, epoch_iter=idx + 1, epoch_size=len(data_loader),
                        time=training_time_meter, is_train=True, loss_c=loss_class_meter, loss_r=loss_weight_meter)
    if pending is not None:
        torch.cuda.synchronize()
        loss_class_meter.update(float(pending[0][0]), pending[1])
        loss_weight_meter.update(float(pending[0][1]), pending[1])
    if writer is not None:
        writer.add_scalar('train/loss_c', loss_class_meter.avg, epoch)
        writer.add_scalar('train/loss_r', loss_weight_meter.avg, epoch)
