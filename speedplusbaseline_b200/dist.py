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
