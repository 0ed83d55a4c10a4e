"""Data parallelism on real GPUs: N ranks, one process per GPU, NCCL.  After one step every rank must hold
IDENTICAL parameters, and they must equal the single-process result on the gradient averaged over the N shards (per-rank
BatchNorm statistics, as in the reference run once per GPU).  Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from oracle import krn as okrn, synth
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
from speedplusbaseline_b200.nets.revgrad import RevGrad
from speedplusbaseline_b200.optim import FusedAdamW, FusedSGD
from speedplusbaseline_b200.core.trainer import KRNTrainStep
from speedplusbaseline_b200.core.dann import DANNTrainStep
from speedplusbaseline_b200 import dist as D
rank, world, local = D.env_world()
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
D.init_process_group('nccl', dev)
mode, out = sys.argv[1], sys.argv[2]
optname = sys.argv[3] if len(sys.argv) > 3 else 'adamw'
B = 4
if mode == 'krn':
    m = KeypointRegressionNet(11, device=dev, seed=100 + rank)       # different init per rank: broadcast must fix it
else:
    m = RevGrad(11, device=dev, seed=100 + rank)
D.broadcast_model(m)
m.train()
if optname == 'sgd':
    opt = FusedSGD(m._store, m.parameters(), lr=0.05, momentum=0.9, weight_decay=0.01, clip_mode=1)
else:
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
x, y = synth.synth_images(B, seed=10 + rank).to(dev), synth.synth_keypoints(B, seed=10 + rank).to(dev)
if mode == 'krn':
    st = KRNTrainStep(m, opt, use_graph=True, world_size=world)
    for _ in range(2):
        st.step(x, y)
else:
    t = synth.synth_images(B, seed=20 + rank).to(dev)
    st = DANNTrainStep(m, opt, use_graph=True, world_size=world)
    for _ in range(2):
        st.step(x, y, t, 0.37)
torch.cuda.synchronize()
p = m._store.params.clone()
ref = p.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(p, ref))
allsame = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
if rank == 0:
    torch.save({'params': p.cpu(), 'identical': int(allsame.item()), 'world': world}, out)
dist.barrier()
dist.destroy_process_group()
'''


def _run(mode, tmp_path, world=2, optname='adamw'):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = str(tmp_path / ('%s.pt' % mode))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', '29611', str(script), mode, out, optname]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    return torch.load(out)


def _single_process_reference(mode, world, optname='adamw'):
    """the same two steps in ONE process: rank r's shard goes through its own forward/backward (own BN statistics), the
    flat gradients are averaged, then one clip + AdamW -- the definition of the N-rank step."""
    from oracle import synth
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.nets.revgrad import RevGrad
    from speedplusbaseline_b200.optim import FusedAdamW, FusedSGD
    dev = torch.device('cuda:0')
    m = (KeypointRegressionNet if mode == 'krn' else RevGrad)(11, device=dev, seed=100)
    m.train()
    if optname == 'sgd':
        opt = FusedSGD(m._store, m.parameters(), lr=0.05, momentum=0.9, weight_decay=0.01, clip_mode=1)
    else:
        opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    opt.grad_scale = 1.0 / world
    eng, st = m.engine, m._store
    bufs0 = st.bufs.clone()
    neg_alpha = torch.full((1,), -0.37, device=dev)
    for _ in range(2):
        st.grads.zero_()
        for r in range(world):
            x, y = synth.synth_images(4, seed=10 + r).to(dev), synth.synth_keypoints(4, seed=10 + r).to(dev)
            if r > 0:
                keep = st.bufs.clone()            # rank 0's BN running statistics are the ones checkpointed; others are per-rank
            if mode == 'krn':
                cx = eng.forward(x, y, train=True)
                eng.backward(cx)
            else:
                t = synth.synth_images(4, seed=20 + r).to(dev)
                cs = eng.forward(x, y, train=True, slot=0)
                eng.domain_forward(cs, 1.0)
                ct = eng.forward(t, None, train=True, slot=1)
                eng.domain_forward(ct, 0.0)
                fs = eng.domain_backward(cs, neg_alpha)
                eng.backward(cs, feature_grad=fs, pose=True)
                ft = eng.domain_backward(ct, neg_alpha)
                eng.backward(ct, feature_grad=ft, pose=False)
            if r > 0:
                st.bufs.copy_(keep)
        opt.step()
    torch.cuda.synchronize()
    return st.params.cpu()


@pytest.mark.parametrize('mode', ['krn', 'dann'])
def test_two_ranks_end_a_step_with_identical_parameters_equal_to_the_shard_average(mode, tmp_path):
    """SGD with momentum: the update is LINEAR in the averaged gradient, so the two-rank result must equal the single-process
    shard average to fp32 summation noise (all-reduce order, atomics)."""
    from kutil import rel
    got = _run(mode, tmp_path, optname='sgd')
    assert got['world'] == 2 and got['identical'] == 1
    ref = _single_process_reference(mode, 2, 'sgd')
    e = rel(got['params'], ref)
    assert e < 2e-4, e


@pytest.mark.parametrize('mode', ['krn', 'dann'])
def test_two_ranks_adamw_step_within_the_optimizer_step_size(mode, tmp_path):
    """AdamW (the reference's optimizer) normalises every gradient component by its own magnitude: the first steps move each
    parameter by ~lr whatever the gradient's size, so fp32 summation noise on the near-zero components is amplified to a fraction
    of lr -- the single-process reference itself differs run to run at batch 4 (tools/gpu_debug_ddp_ref.py: median 1e-7 or 7e-5,
    max 1.6e-3 .. 3.5e-3 after two steps at lr 1e-3).  What data parallelism must guarantee survives that: identical parameters on
    every rank, no parameter further from the reference than the two steps can move it, and the bulk within a tenth of a step."""
    got = _run(mode, tmp_path, optname='adamw')
    assert got['world'] == 2 and got['identical'] == 1
    ref = _single_process_reference(mode, 2, 'adamw')
    d = (got['params'].double() - ref.double()).abs()
    lr, steps = 1e-3, 2
    assert float(d.max()) <= 2.2 * steps * lr, float(d.max())
    assert float(d.median()) < 0.2 * lr, float(d.median())
