"""N>1 host logic on CPU: world_size-2 gloo group.  The compute on each rank is the
oracle (test infrastructure) -- what is under test is the data-parallel plumbing of
speedplusbaseline_b200.dist: shard -> per-rank gradients (per-rank BatchNorm statistics) -> flat-buffer
sum-allreduce -> 1/world scale -> clip on the averaged gradient -> identical AdamW on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flatten(keys, d):
    return torch.cat([d[k].reshape(-1) for k in keys])


def _rank_step(sd, keys, x, y, sync):
    """one data-parallel KRN step on this rank's shard; returns the flat averaged gradient"""
    from oracle import krn as okrn
    from oracle.optim import AdamWState, adamw_step, clip_grad_norm
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    loss, _ = okrn.krn_forward(sd, x, y, train=True)
    loss.backward()
    grads = {k: (sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])) for k in keys}
    for k in keys:
        sd[k].requires_grad_(False)
    flat = _flatten(keys, grads).contiguous()
    sync.allreduce(flat)
    flat *= sync.grad_scale
    return flat


def _apply(sd, keys, flat):
    from oracle.optim import AdamWState, adamw_step, clip_grad_norm
    gl, off = [], 0
    for k in keys:
        n = sd[k].numel()
        gl.append(flat[off:off + n].view_as(sd[k]).clone())
        off += n
    total = clip_grad_norm(gl, 1.0)
    st = AdamWState([sd[k] for k in keys])
    adamw_step([sd[k] for k in keys], gl, st, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01)
    return float(total)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from oracle import krn as okrn, synth, steps
    from speedplusbaseline_b200.dist import GradSync, init_process_group, shard
    r, w, _ = init_process_group('gloo')
    assert (r, w) == (rank, world)
    # --- bucketed == unbucketed == analytic sum
    base = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    a, b = base.clone(), base.clone()
    GradSync(w).allreduce(a)
    GradSync(w, bucket_elems=300).allreduce(b)
    want = torch.arange(1000, dtype=torch.float32) * sum(range(1, w + 1))
    assert torch.equal(a, want) and torch.equal(b, want)
    assert GradSync(w, bucket_elems=300).buckets(1000) == [(700, 1000), (400, 700), (100, 400), (0, 100)]
    # --- one data-parallel KRN step, per-GPU batch 1, global batch 2
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    keys = steps.param_keys(sd)
    X, Y = synth.synth_images(w), synth.synth_keypoints(w)
    sync = GradSync(w)
    flat = _rank_step(sd, keys, shard(X, r, w), shard(Y, r, w), sync)
    total = _apply(sd, keys, flat)
    chk = torch.tensor([synth.checksum(sd['head.0.weight']), synth.checksum(sd['base.1.conv.0.0.weight']), total], dtype=torch.float64)
    allc = [torch.zeros_like(chk) for _ in range(w)]
    dist.all_gather(allc, chk)
    for c in allc:                              # every replica holds bit-identical parameters after the step
        assert torch.equal(c, allc[0])
    if rank == 0:
        # single-process emulation: the oracle on each shard in turn, gradients averaged
        sd1 = synth.synth_state_dict(okrn.krn_shapes(), 2021)
        acc = None
        for rr in range(w):
            sdr = {k: v.clone() for k, v in sd1.items()}
            f = _rank_step(sdr, keys, shard(X, rr, w), shard(Y, rr, w), GradSync(1))
            acc = f if acc is None else acc + f
        acc /= w
        total1 = _apply(sd1, keys, acc)
        assert abs(total1 - total) <= 1e-5 * abs(total1)
        err = float((sd1['head.0.weight'] - sd['head.0.weight']).abs().max())
        assert err <= 1e-6, err
        q.put('ok')
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_grad_allreduce_step():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == 'ok'
