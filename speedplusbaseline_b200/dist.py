"""Data-parallel plumbing: one process per GPU (torchrun), batch-sharded replicas, ONE exchange per
step -- a sum-allreduce of the flat fp32 gradient buffer (the reference itself is
single-GPU, train.py:50).  The 1/world factor is not applied here: it is folded into the fused
AdamW kernel's `grad_scale`, and the clip norm is taken on the averaged gradient, identical on every
rank, so there is no second collective.  Backend: NCCL over NVLink on GPUs, gloo for the CPU tests."""
import os

import torch
import torch.distributed as dist


def env_world():
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0')))


def init_process_group(backend=None, device=None):
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {}
        if backend == 'nccl' and device is not None:
            kw['device_id'] = device
        dist.init_process_group(backend, **kw)
    return rank, world, local


class GradSync:
    """Sum-allreduce of a flat gradient buffer, optionally in buckets (reverse order = the order the
    backward pass finishes them, so a bucket can be sent while earlier layers still compute)."""

    def __init__(self, world_size=1, group=None, bucket_elems=0):
        self.world, self.group, self.bucket = int(world_size), group, int(bucket_elems)

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def buckets(self, n):
        if self.bucket <= 0 or self.bucket >= n:
            return [(0, n)]
        out, hi = [], n
        while hi > 0:
            lo = max(0, hi - self.bucket)
            out.append((lo, hi))
            hi = lo
        return out

    def allreduce(self, flat, async_op=False):
        if self.world <= 1:
            return []
        works = []
        for lo, hi in self.buckets(flat.numel()):
            w = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                works.append(w)
        return works


def shard(t, rank, world):
    """contiguous batch shard of a global-batch tensor (per-GPU batch = global / world)."""
    n = t.shape[0] // world
    return t[rank * n:(rank + 1) * n]


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def setup_cli(cfg):
    """Device + process group for train.py / adapt.py / test.py.  Single process: the reference's own rule
    (cuda:0 when available and not --no_cuda, train.py:50).  Under torchrun (WORLD_SIZE > 1): one process per GPU,
    cuda:LOCAL_RANK, NCCL process group; `cfg.device` tells get_model where to build the engine."""
    r, world, local = env_world()
    use_cuda = torch.cuda.is_available() and cfg.use_cuda
    if world > 1:
        if use_cuda:
            torch.cuda.set_device(local)
            device = torch.device('cuda', local)
            init_process_group('nccl', device)
        else:
            device = torch.device('cpu')
            init_process_group('gloo')
    else:
        device = torch.device('cuda:0') if use_cuda else torch.device('cpu')
    cfg.device = str(device) if device.type == 'cuda' else None      # a string: cfg.__dict__ is dumped to config.txt as JSON
    cfg.rank, cfg.world_size = r, world
    return device


def broadcast_model(model):
    """Replicas must start from identical parameters / BN buffers (every rank constructs its model with its own RNG
    state): rank 0's flat buffers are broadcast once.  No-op for a single process."""
    if world_size() <= 1:
        return
    st = getattr(model, '_store', None)
    if st is not None:
        for t in (st.params, st.bufs, st.nbt):
            dist.broadcast(t, 0)
        if getattr(st, 'params_lowp', None) is not None:
            st.params_lowp.copy_(st.params)
    else:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)


def is_main():
    return rank() == 0
