"""`--use_fp16` on this path = fp32 storage + SINGLE-pass TF32 tensor-core GEMMs (include/b200sp.h B200SP_F32_TF32X1): operands
carry fp16's 10-bit mantissa, the exponent range and the accumulation stay fp32.  The reference's flag runs torch autocast
(trainer.py:73-94); the yardstick here is therefore the reference algorithm under torch.autocast on the CPU (bfloat16, the
only CPU autocast type with conv support: 8-bit mantissa, i.e. COARSER than TF32), both measured against float64:
the TF32 path must be at least as close to float64 as autocast is."""
import pytest
import torch

from oracle import krn as okrn, synth, steps
from kutil import rel

pytestmark = pytest.mark.gpu


def _model(sd, tf32=True):
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    m = KeypointRegressionNet(11, device='cuda:0', seed=1, tf32_gemm=tf32)
    m.load_state_dict(sd)
    return m


def _oracle(sd, x, y, dt, autocast=False):
    s = {k: (v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    st = steps.new_state(s)
    if autocast:
        with torch.autocast('cpu', dtype=torch.bfloat16):
            return steps.krn_train_step(s, st, x.to(dt), y.to(dt)), s
    return steps.krn_train_step(s, st, x.to(dt), y.to(dt)), s


def test_eval_logits_at_tf32_resolution():
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    x = synth.synth_images(4)
    sdo = {k: v.clone().double() if v.is_floating_point() else v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ref = okrn.krn_logits(sdo, x.double(), False)[1]
    m = _model(sd).eval()
    xc, yc = m(x.cuda())
    got = torch.stack([xc, yc], 2).reshape(4, 22).double()
    e_tf32 = rel(got, ref)
    m32 = _model(sd, tf32=False).eval()
    xc, yc = m32(x.cuda())
    e_f32 = rel(torch.stack([xc, yc], 2).reshape(4, 22).double(), ref)
    print('eval logits rel-L2 vs float64: single-pass TF32 %.2e   3xTF32 %.2e' % (e_tf32, e_f32))
    # measured on B200 with the synthetic weights: 5.8e-5 (3xTF32) vs 4.5e-2 (single pass) -- the SURVEY's finding that one
    # TF32 pass is NOT fp32 parity; it is the mixed-precision mode, and the training test below pins it against autocast
    assert e_f32 < 1e-4 and 1e-4 < e_tf32 < 1e-1


def test_train_step_no_worse_than_autocast_reference():
    from speedplusbaseline_b200.optim import FusedAdamW
    B = 8
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    x, y = synth.synth_images(B), synth.synth_keypoints(B)
    r64, _ = _oracle(sd, x, y, torch.float64)
    rac, _ = _oracle(sd, x, y, torch.float32, autocast=True)
    m = _model(sd).train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    opt.zero_grad()
    loss, _ = m(x.cuda(), y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    e_loss, e_loss_ac = abs(float(loss) - r64['loss']) / abs(r64['loss']), abs(rac['loss'] - r64['loss']) / abs(r64['loss'])
    gd = m.grad_dict()
    gn = r64['grad_norm']
    worse, n, se, sa = 0, 0, 0.0, 0.0
    for k, g64 in r64['grads'].items():
        if float(g64.norm()) < 1e-3 * gn:
            continue
        e, ea = rel(gd[k], g64), rel(rac['grads'][k], g64)
        n, se, sa = n + 1, se + e, sa + ea
        worse += e > 1.5 * ea + 1e-3
    print('train step B=%d: loss rel err TF32 %.2e / autocast-bf16 %.2e; mean gradient rel-L2 TF32 %.2e / autocast %.2e; '
          '%d of %d tensors worse than 1.5x autocast' % (B, e_loss, e_loss_ac, se / n, sa / n, worse, n))
    assert e_loss <= max(2.0 * e_loss_ac, 2e-3)
    assert se / n <= sa / n and worse <= n // 20
    assert all(torch.isfinite(v).all() for v in gd.values())


def test_epoch_loop_accepts_the_grad_scaler_and_trains(tmp_path):
    """train.py:102-103 creates a GradScaler for --use_fp16 and hands it to the epoch loop: accepted, scale untouched
    (fp32 exponent range: nothing to scale); the captured step stays finite and tracks the 3xTF32 path's first steps."""
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    m = _model(sd).train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    stp = KRNTrainStep(m, opt, use_graph=True)
    m32 = _model(sd, tf32=False).train()
    opt32 = FusedAdamW(m32._store, m32.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    stp32 = KRNTrainStep(m32, opt32, use_graph=True)
    x, y = synth.synth_images(8).cuda(), synth.synth_keypoints(8).cuda()
    losses = [float(stp.step(x, y)[0]) for _ in range(4)]
    losses32 = [float(stp32.step(x, y)[0]) for _ in range(4)]
    assert all(l == l and l < 1e9 for l in losses), losses
    # random-init KRN at lr 1e-3 is chaotic after the first update (the 3xTF32 path's own losses jump by 100x): the first two
    # steps must agree with the full-precision path to mixed-precision accuracy
    assert abs(losses[0] - losses32[0]) <= 2e-3 * abs(losses32[0]), (losses, losses32)
    assert abs(losses[1] - losses32[1]) <= 0.2 * abs(losses32[1]), (losses, losses32)
    scaler = torch.amp.GradScaler('cuda')
    from speedplusbaseline_b200.core.trainer import train_single_epoch_krn
    import types
    cfg = types.SimpleNamespace(texture_ratio=0.0, use_graph=True)
    train_single_epoch_krn(1, cfg, m, [(x.cpu().pin_memory(), y.cpu().pin_memory())] * 2, opt, None, torch.device('cuda:0'), scaler=scaler)
    assert scaler.get_scale() == 65536.0            # untouched: nothing to scale in an fp32-range path
