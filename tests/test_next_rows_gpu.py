"""SURVEY.md 8(f) rows on the GPU, through the C-ABI: fused SGD / RMSprop / Adam over the flat buffer against the
oracle (float64) and against the fixtures the reference's own get_optimizer produced (tests/golden/optim_steps.npz)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from speedplusbaseline_b200 import _lib as L
from oracle import optim as ooptim
from oracle.make_golden_optim import HYPER, STEPS, synth_problem
from kutil import rel, sp

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _hp_block(**kw):
    h = L.AdamWHp()
    for k, v in kw.items():
        setattr(h, k, v)
    return torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).cuda()


KINDS = {'adamw': L.OPT_ADAMW, 'sgd': L.OPT_SGD, 'rmsprop': L.OPT_RMSPROP, 'adam': L.OPT_ADAM}


@pytest.mark.parametrize('name', ['sgd', 'rmsprop', 'adam', 'adamw'])
def test_optim_step_matches_reference_factory_golden(golden_dir, name):
    """raw C-ABI call on a 1031-element buffer (vector body + scalar tail), clip_grad_norm_(1.0) folded in."""
    gold = np.load(os.path.join(golden_dir, 'optim_steps.npz'))[name]
    p0, grads = synth_problem()
    p = torch.cat([t.reshape(-1) for t in p0]).cuda()
    s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
    hp = _hp_block(lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, weight_decay=HYPER['weight_decay'],
                   max_norm=1.0, clip_value=1.0, grad_scale=1.0, step=0, clip_mode=1)
    for s in range(STEPS):
        g = torch.cat([t.reshape(-1) for t in grads[s]]).cuda()
        L.call('b200sp_grad_sqnorm', g.data_ptr(), p.numel(), hp.data_ptr(), sp())
        L.call('b200sp_optim_step', KINDS[name], p.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, p.numel(),
               hp.data_ptr(), sp())
        torch.cuda.synchronize()
        np.testing.assert_allclose(p.cpu().numpy(), gold[s], rtol=3e-6, atol=3e-7, err_msg='%s step %d' % (name, s))
    h = L.AdamWHp.from_buffer_copy(hp.cpu().numpy().tobytes())
    assert h.step == STEPS and h.sqnorm == 0.0


@pytest.mark.parametrize('name,clip_mode', [('sgd', 0), ('sgd', 1), ('rmsprop', 1), ('rmsprop', 2), ('adam', 0), ('adam', 2)])
def test_fused_optimizer_classes_match_oracle(name, clip_mode):
    from speedplusbaseline_b200.params import ParamStore
    from speedplusbaseline_b200 import optim as fo
    g = _g(31 + clip_mode + len(name))
    st = ParamStore([('a.weight', 'plain', (37, 5)), ('b.weight', 'plain', (1001,))], [('bn', 8)], torch.device('cuda'))
    p0 = torch.randn(st.n, generator=g)
    st.params.copy_(p0)
    par = [torch.nn.Parameter(st.params)]
    if name == 'sgd':
        opt = fo.FusedSGD(st, par, lr=1e-2, momentum=0.9, weight_decay=0.01, clip_mode=clip_mode)
    elif name == 'rmsprop':
        opt = fo.FusedRMSprop(st, par, lr=1e-3, alpha=0.9, weight_decay=0.01, clip_mode=clip_mode)
    else:
        opt = fo.FusedAdam(st, par, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=clip_mode)
    ref_p = [p0.clone().double()]
    ost = ooptim.FlatState(ref_p)
    for it in range(4):
        gr = torch.randn(st.n, generator=g) * (3.0 if it == 1 else 0.01)
        st.grads.copy_(gr)
        opt.step()
        gl = [gr.clone().double()]
        if clip_mode == 1:
            ooptim.clip_grad_norm(gl, 1.0)
        elif clip_mode == 2:
            ooptim.clip_grad_value(gl, 1.0)
        if name == 'sgd':
            ooptim.sgd_step(ref_p, gl, ost, lr=1e-2, momentum=0.9, wd=0.01)
        elif name == 'rmsprop':
            ooptim.rmsprop_step(ref_p, gl, ost, lr=1e-3, alpha=0.9, eps=1e-8, wd=0.01)
        else:
            ooptim.adam_step(ref_p, gl, ost, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01)
        torch.cuda.synchronize()
        assert rel(st.params, ref_p[0]) < 2e-6, it
    sd = opt.state_dict()                      # torch's own checkpoint layout: per-parameter state in the reference shapes
    n1 = {'sgd': 'momentum_buffer', 'rmsprop': 'square_avg', 'adam': 'exp_avg'}[name]
    assert len(sd['state']) == 4 and sd['param_groups'][0]['params'] == [0, 1, 2, 3]       # a.weight, b.weight, bn.weight, bn.bias
    assert tuple(sd['state'][0][n1].shape) == (37, 5) and tuple(sd['state'][3][n1].shape) == (8,)
    assert ('exp_avg_sq' in sd['state'][0]) == (name == 'adam')
    if name != 'sgd':
        assert float(sd['state'][1]['step']) == 4.0
    assert opt.device_step() == 4
    assert rel(opt.exp_avg, ost.s1[0]) < 1e-5


def test_optim_step_rejects_bad_arguments():
    p = torch.zeros(8, device='cuda')
    hp = _hp_block(lr=1e-3)
    assert L.lib.b200sp_optim_step(7, p.data_ptr(), p.data_ptr(), p.data_ptr(), None, None, 8, hp.data_ptr(), sp()) == -22
    assert L.lib.b200sp_optim_step(L.OPT_ADAM, p.data_ptr(), p.data_ptr(), p.data_ptr(), None, None, 8, hp.data_ptr(), sp()) == -22


# ---- row f2: evaluation tail on the device --------------------------------------------------------
def _tie_free_rows(w, k):
    s = -np.sort(-w, axis=1)[:, :k + 1]
    return np.array([len(np.unique(r)) == k + 1 for r in s])


def test_topk_softmax_matches_reference_loop_golden(golden_dir):
    from oracle.make_golden_postproc import K_NB, synth_inputs
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    w = synth_inputs()[0]
    B, N = w.shape
    wd = w.cuda()
    tw, traw = torch.empty(B, K_NB, device='cuda'), torch.empty(B, K_NB, device='cuda')
    ti = torch.empty(B, K_NB, dtype=torch.int64, device='cuda')
    L.call('b200sp_topk_softmax', wd.data_ptr(), tw.data_ptr(), traw.data_ptr(), ti.data_ptr(), B, N, K_NB, sp())
    torch.cuda.synchronize()
    ti_h, free = ti.cpu().numpy(), _tie_free_rows(w.numpy(), K_NB)
    assert (ti_h[free] == g['top_idx'][free]).all()                                         # bit-exact indices
    assert (np.take_along_axis(w.numpy(), ti_h, 1) == np.take_along_axis(w.numpy(), g['top_idx'], 1)).all()
    assert (traw.cpu().numpy() == np.take_along_axis(w.numpy(), ti_h, 1)).all()
    # softmax: expf on the device vs torch's vectorised exp on the host -- 1e-6 relative (written tolerance)
    np.testing.assert_allclose(tw.cpu().numpy(), g['top_w'], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('B,N,k', [(1, 5, 5), (3, 33, 1), (32, 5000, 5), (7, 50001, 32), (2, 256, 8)])
def test_topk_softmax_matches_oracle(B, N, k):
    from oracle import postproc
    w = torch.randn(B, N, generator=_g(B + N + k)) * 2
    w[0, N // 2] = float('-inf')
    wd = w.cuda()
    tw, ti = torch.empty(B, k, device='cuda'), torch.empty(B, k, dtype=torch.int64, device='cuda')
    L.call('b200sp_topk_softmax', wd.data_ptr(), tw.data_ptr(), None, ti.data_ptr(), B, N, k, sp())
    torch.cuda.synchronize()
    ow, oi = postproc.spn_top_classes(w.numpy(), k)
    assert (ti.cpu().numpy() == oi).all()
    np.testing.assert_allclose(tw.cpu().numpy(), ow, rtol=1e-6, atol=1e-30)
    tv, tix = torch.topk(w, k, dim=1)                               # and the torch ops the reference calls
    assert (tix.numpy() == ti.cpu().numpy()).all()
    np.testing.assert_allclose(tw.cpu().numpy(), torch.softmax(tv, 1).numpy(), rtol=1e-6, atol=1e-30)


def test_topk_softmax_argument_errors():
    w = torch.zeros(2, 8, device='cuda')
    o, i = torch.zeros(2, 9, device='cuda'), torch.zeros(2, 9, dtype=torch.int64, device='cuda')
    assert L.lib.b200sp_topk_softmax(w.data_ptr(), o.data_ptr(), None, i.data_ptr(), 2, 8, 9, sp()) == -22      # k > N
    assert L.lib.b200sp_topk_softmax(w.data_ptr(), o.data_ptr(), None, i.data_ptr(), 2, 8, 0, sp()) == -22
    assert L.lib.b200sp_topk_softmax(w.data_ptr(), o.data_ptr(), None, i.data_ptr(), 0, 8, 2, sp()) == 0        # empty batch


def test_kpt_denorm_bit_exact_vs_reference_golden(golden_dir):
    from oracle.make_golden_postproc import synth_inputs
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    _, x, y, bb = synth_inputs()
    B, K = x.shape
    logits = torch.stack([x, y], dim=2).reshape(B, 2 * K).contiguous().cuda()      # interleaved like the head output
    out = torch.empty(B, K, 2, device='cuda')
    L.call('b200sp_kpt_denorm', logits.data_ptr(), bb.cuda().data_ptr(), out.data_ptr(), B, K, sp())
    torch.cuda.synchronize()
    assert (out.cpu().numpy() == g['kpt_pix']).all()


class _Loader(list):
    pass


def _stub_tail(record):
    """CPU pose code replaced by recorders (EPnP / metrics are out of scope and need the reference checkout)."""
    def pnp(c3, pix, cam, dist):
        record.setdefault('pix', []).append(np.array(pix))
        return np.array([1.0, 0, 0, 0]), np.zeros(3)

    def wmq(qs, w):
        record.setdefault('qs', []).append(np.array(qs))
        record.setdefault('w', []).append(np.array(w))
        return np.array([1.0, 0, 0, 0])
    return dict(pnp=pnp, weighted_mean_quaternion=wmq, compute_position_spn=lambda *a: np.zeros(3),
                error_orientation=lambda a, b: 1.5, error_translation=lambda a, b: 0.25,
                speed_score=lambda *a, **k: (0.5, 1.0))


def test_valid_krn_batched_matches_per_image_reference_flow(tmp_path):
    """batch-4 evaluation through core.inference.valid_krn == the reference flow (batch-1 forward, .cpu(), numpy
    de-normalisation) on the same model: keypoint pixels agree to fp32 round-off of the logits."""
    from types import SimpleNamespace
    from oracle import postproc
    from speedplusbaseline_b200.core import inference
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    m = KeypointRegressionNet(11, device='cuda', seed=5)
    g = _g(3)
    imgs = torch.rand(4, 3, 224, 224, generator=g)
    bbox = torch.tensor([[10., 500., 20., 480.], [0., 1920., 0., 1200.], [100.5, 300.25, 50., 400.], [5., 6., 7., 9.]])
    loader = _Loader([(imgs, bbox, torch.zeros(4, 4), torch.zeros(4, 3))])
    rec = {}
    cfg = SimpleNamespace(logdir=str(tmp_path))
    perf = inference.valid_krn(0, cfg, m, loader, None, None, None, None, torch.device('cuda'), ref=_stub_tail(rec))
    assert perf['eR'].avg == 1.5 and perf['eT'].avg == 0.25 and len(rec['pix']) == 4
    assert open(os.path.join(str(tmp_path), 'err_q.txt')).read().split() == ['1.50000'] * 4
    m.eval()
    with torch.no_grad():
        xc, yc = m(imgs.cuda())                          # the module's reference contract: (xc.cpu(), yc.cpu())
    ref_pix = postproc.krn_keypoints_pix(xc.numpy(), yc.numpy(), bbox.numpy())
    # the head FC accumulates split-K partial sums with fp32 atomics, so two forwards of the same batch agree to fp32
    # round-off, not bit for bit (the de-normalisation itself is bit-exact: test_kpt_denorm_bit_exact_vs_reference_golden)
    np.testing.assert_allclose(np.stack(rec['pix']), ref_pix, rtol=1e-5, atol=1e-3)


def test_valid_spn_batched_top_classes():
    from types import SimpleNamespace
    from oracle import postproc
    from speedplusbaseline_b200.core import inference
    from speedplusbaseline_b200.nets.spn import SpacecraftPoseNet
    m = SpacecraftPoseNet(5000, pretrain=False, device='cuda', seed=9)
    st = m._store
    st.params.normal_(0.0, 0.02, generator=torch.Generator(device='cuda').manual_seed(1))
    imgs = torch.rand(3, 3, 227, 227, generator=_g(4))
    qClass = np.arange(5000 * 4, dtype=np.float64).reshape(5000, 4)
    loader = _Loader([(imgs, torch.zeros(3, 4), torch.zeros(3, 4), torch.zeros(3, 3))])
    rec = {}
    cfg = SimpleNamespace(num_neighbors=5)
    perf = inference.valid_spn(0, cfg, m, loader, None, None, None, None, torch.device('cuda'), qClass, ref=_stub_tail(rec))
    assert perf['speed (raw)'].avg == 0.5
    m.eval()
    with torch.no_grad():
        _, r = m(imgs.cuda())
    ow, oi = postproc.spn_top_classes(r.cpu().numpy(), 5)
    assert (np.stack(rec['qs'])[:, :, 0] / 4 == oi).all()
    np.testing.assert_allclose(np.stack(rec['w']), ow, rtol=1e-4)      # two forwards: split-K fp32 atomics reorder


# ---- row f1: input pipeline on the device -----------------------------------------------------------
def _aug(box, dec=None, std=0.0, seed=1):
    dec = dec or dict(rot=0, flip=0, bc=None)
    bc = dec.get('bc')
    return L.Aug(box[0], box[1], box[2], box[3], dec['rot'], dec['flip'], 1 if bc is not None else 0,
                 float(bc[0]) if bc is not None else 1.0, float(bc[1]) if bc is not None else 0.0, std, seed, 0)


def _replay(case):
    """decisions + expected tensors of one golden case, replayed through the oracle (test_next_rows_cpu pins this replay
    to the reference bit for bit)."""
    from oracle import transforms as ot
    from oracle.make_golden_transforms import FRAME_HW, synth_frame, synth_keypoints
    seed, bbox, p, is_train, model = case
    H, W = FRAME_HW
    grey = synth_frame(seed)
    kp = synth_keypoints(seed, bbox)
    gen = torch.Generator().manual_seed(seed)
    size = (224, 224) if model == 'krn' else (227, 227)
    dec = dict(rot=0, flip=0, bc=None, noise=None)
    if model == 'krn':
        u = [torch.rand(1, generator=gen) for _ in range(3)] if is_train else None
        box = ot.random_crop_box(bbox, W, H, is_train, u)
        k = ot.crop_keypoints(kp, box)
    else:
        box = ot.resize_crop_box(bbox, W, H)
        k = torch.as_tensor(kp)
    img = ot.crop_resize_to_tensor(np.repeat(grey[:, :, None], 3, 2), box, size)
    if is_train and model == 'krn':
        dec = ot.draw_augment(p, img.shape, gen)
    nonoise = {k2: v for k2, v in dec.items() if k2 != 'noise'}
    img_nn, k = ot.apply_augment(img, k, **nonoise)
    return grey, kp, box, dec, img_nn, k


@pytest.mark.parametrize('channels', [1, 3])
def test_input_pipeline_bit_exact_vs_reference_transforms(golden_dir, channels):
    from oracle.make_golden_transforms import CASES
    from speedplusbaseline_b200.datasets.transforms import DeviceTransforms
    g = np.load(os.path.join(golden_dir, 'transforms.npz'))
    for model, size in (('krn', (224, 224)), ('spn', (227, 227))):
        cases = [c for c in CASES if c[4] == model]
        rp = [_replay(c) for c in cases]
        frames = torch.from_numpy(np.stack([r[0] for r in rp]))
        if channels == 3:
            frames = frames.unsqueeze(-1).repeat(1, 1, 1, 3)
        tf = DeviceTransforms(model, size, device='cuda')
        augs = [_aug(r[2], r[3]) for r in rp]
        out, kout = tf.apply(frames, augs, np.stack([r[1] for r in rp]), normalize_kpts=model == 'krn')
        torch.cuda.synchronize()
        assert tf.status() == 0
        for i, (c, r) in enumerate(zip(cases, rp)):
            assert torch.equal(out[i].cpu(), r[4]), ('image', c[0])                    # vs the oracle (noise switched off)
            assert torch.equal(kout[i].cpu(), r[5]), ('keypoints', c[0])
            assert np.array_equal(kout[i].cpu().numpy(), g['kpt%d' % c[0]])
            if r[3]['noise'] is None:                                                    # no noise drawn: the reference output itself
                assert np.array_equal(out[i].cpu().numpy(), g['img%d' % c[0]]), ('golden', c[0])


def test_input_pipeline_full_size_frames_vs_oracle():
    """SPEED+ geometry: 1200x1920 grey frames, RoIs from tiny to the whole frame height, every rotation / flip."""
    from oracle import transforms as ot
    from speedplusbaseline_b200.datasets.transforms import DeviceTransforms
    rng = np.random.default_rng(11)
    H, W, B = 1200, 1920, 6
    frames = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    boxes = [(0, 1200, 0, 1200), (700, 1920, 0, 1200), (900, 1000, 500, 601), (13, 237, 977, 1200), (300, 1500, 100, 1150), (1000, 1224, 400, 624)]
    decs = [dict(rot=i % 4, flip=i % 3, bc=(0.5 + 0.3 * i, 0.05 * (i - 2)) if i % 2 else None) for i in range(B)]
    tf = DeviceTransforms('krn', (224, 224), device='cuda')
    out, _ = tf.apply(torch.from_numpy(frames), [_aug(b, d) for b, d in zip(boxes, decs)])
    torch.cuda.synchronize()
    assert tf.status() == 0
    for i in range(B):
        img = ot.crop_resize_to_tensor(np.repeat(frames[i][:, :, None], 3, 2), boxes[i], (224, 224))
        img, _ = ot.apply_augment(img, torch.zeros(2, 1), **decs[i])
        assert torch.equal(out[i].cpu(), img), i


def test_input_pipeline_noise_statistics():
    """GaussianNoise :101-112 is N(0, (25/255)^2) per element; the device generator is checked statistically."""
    from speedplusbaseline_b200.datasets.transforms import DeviceTransforms
    frames = torch.full((2, 300, 300), 128, dtype=torch.uint8)
    tf = DeviceTransforms('krn', (224, 224), device='cuda')
    std = 25 / 255
    a1, _ = tf.apply(frames, [_aug((0, 300, 0, 300), std=std, seed=5), _aug((0, 300, 0, 300), std=std, seed=6)])
    a2, _ = tf.apply(frames, [_aug((0, 300, 0, 300), std=std, seed=5), _aug((0, 300, 0, 300), std=0.0)])
    torch.cuda.synchronize()
    base = 128 / 255
    assert torch.equal(a2[1].cpu(), torch.full((3, 224, 224), base))
    assert torch.equal(a1[0], a2[0])                                   # same seed, same noise
    n = (a1.double().cpu() - base)                                      # 0.5 +- 5 sigma stays inside [0,1]: no clamping
    N = n[0].numel()
    for i in range(2):
        assert abs(float(n[i].mean())) < 5 * std / N ** 0.5
        assert abs(float(n[i].std()) / std - 1) < 0.01
        assert abs(float((n[i] ** 4).mean()) / std ** 4 - 3) < 0.1      # Gaussian kurtosis
    c = np.corrcoef(n[0].reshape(3, -1).numpy())
    assert abs(c[0, 1]) < 0.02 and abs(c[1, 2]) < 0.02                  # channels get independent noise
    assert abs(np.corrcoef(n[0].reshape(-1).numpy(), n[1].reshape(-1).numpy())[0, 1]) < 0.02    # and so do seeds


def test_input_pipeline_guards_and_errors():
    frames = torch.randint(0, 256, (1, 64, 64, 1), dtype=torch.uint8, device='cuda')
    aug = torch.frombuffer(bytearray(bytes((L.Aug * 1)(L.Aug(0, 64, 0, 64, 0, 0, 0, 1.0, 0.0, 0.0, 1, 0)))), dtype=torch.uint8).cuda()
    oh = ow = 8                                       # scale 8 -> 17 taps needed
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    out = torch.full((1, 3, oh, ow), -1.0, device='cuda')

    def run(ks, tmp_rows, C_=1, odd=0, oh_=oh):
        coef = torch.zeros(2 * 8 * (2 + ks), dtype=torch.int32, device='cuda')
        tmp = torch.zeros(max(1, tmp_rows) * ow, dtype=torch.uint8, device='cuda')
        return L.lib.b200sp_input_pipeline(frames.data_ptr(), 1, 64, 64, C_, aug.data_ptr(), coef.data_ptr(), tmp.data_ptr(), tmp_rows,
                                           ks, out.data_ptr(), oh_, ow, odd, status.data_ptr(), sp())
    assert run(17, 64) == 0
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and float(out.min()) >= 0.0
    assert run(5, 64) == 0                            # too few taps: flagged, truncated, in bounds
    torch.cuda.synchronize()
    assert int(status.item()) & 2
    status.zero_()
    assert run(17, 32) == 0                           # temp too small for the crop: flagged, zeros
    torch.cuda.synchronize()
    assert int(status.item()) & 1 and float(out.abs().max()) == 0.0
    assert run(17, 64, C_=2) == -22 and run(2, 64) == -22 and run(17, 64, odd=1, oh_=4) == -22
    from speedplusbaseline_b200.datasets.transforms import DeviceTransforms
    tf = DeviceTransforms('krn', (8, 8), device='cuda')
    with pytest.raises(ValueError):
        tf.apply(frames, [L.Aug(0, 65, 0, 64, 0, 0, 0, 1.0, 0.0, 0.0, 1, 0)])
    with pytest.raises(TypeError):
        tf.apply(frames.float(), [L.Aug(0, 64, 0, 64, 0, 0, 0, 1.0, 0.0, 0.0, 1, 0)])


def test_device_transforms_feed_the_training_step():
    """build_transforms(...) output is what the fused KRN step consumes: frames -> images/keypoints -> one train step."""
    from speedplusbaseline_b200.datasets.transforms import build_transforms
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    rng = np.random.default_rng(2)
    B, H, W = 4, 600, 960
    frames = torch.from_numpy(rng.integers(0, 256, (B, H, W), dtype=np.uint8)).pin_memory()
    bbox = np.array([[100, 400, 50, 300], [300, 900, 100, 580], [10, 200, 20, 120], [500, 800, 200, 590]], np.float32)
    kp = np.stack([np.stack([rng.uniform(b[0], b[1], 11), rng.uniform(b[2], b[3], 11)]) for b in bbox]).astype(np.float32)
    tf = build_transforms('krn', (224, 224), p_aug=0.5, is_train=True, device='cuda', generator=torch.Generator().manual_seed(3))
    images, bb, kpts = tf(frames, bbox, kp)
    torch.cuda.synchronize()
    assert tf.status() == 0 and images.shape == (B, 3, 224, 224) and kpts.shape == (B, 2, 11) and bb.shape == (B, 4)
    assert float(images.min()) >= 0.0 and float(images.max()) <= 1.0 and float(images.std()) > 0.05
    m = KeypointRegressionNet(11, device='cuda', seed=1)
    m.train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    loss3 = KRNTrainStep(m, opt, use_graph=False).step(images, kpts.contiguous())
    torch.cuda.synchronize()
    assert torch.isfinite(loss3).all() and float(loss3[0]) > 0


def test_device_batch_loader_drives_the_reference_epoch_loop(tmp_path):
    """decode-only dataset -> DeviceBatchLoader -> train_single_epoch_krn (unchanged loop, DevicePrefetcher included)."""
    from types import SimpleNamespace
    from test_next_rows_cpu import _write_split
    from speedplusbaseline_b200.datasets.raw import make_dataloader
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.trainer import train_single_epoch_krn
    _write_split(str(tmp_path), 8, hw=(300, 400))
    cfg = SimpleNamespace(dataroot=str(tmp_path), dataname='speedplus', num_keypoints=11, model_name='krn', train_domain='synthetic',
                          test_domain='lightbox', train_csv='train.csv', test_csv='lightbox.csv', batch_size=4, num_workers=0,
                          input_shape=(224, 224), texture_ratio=0.5, use_graph=False)
    dev = torch.device('cuda')
    loader = make_dataloader(cfg, is_train=True, is_source=True, load_labels=True, device=dev, generator=_g(1))
    assert len(loader) == 2
    m = KeypointRegressionNet(11, device='cuda', seed=1)
    p0 = m._store.params.clone()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    train_single_epoch_krn(1, cfg, m, loader, opt, None, dev)
    torch.cuda.synchronize()
    assert loader.tf.status() == 0
    assert torch.isfinite(m._store.params).all() and float((m._store.params - p0).abs().max()) > 0
    assert int(m._store.nbt[0]) == 2
    test_loader = make_dataloader(cfg, is_train=False, is_source=False, load_labels=True, device=dev)
    images, bbox, q, t = next(iter(test_loader))
    assert images.shape == (1, 3, 224, 224) and bbox.shape == (1, 4) and q.shape == (1, 4) and t.shape == (1, 3)


@pytest.mark.parametrize('name', ['sgd', 'rmsprop', 'adam'])
def test_fused_optimizers_inside_the_captured_krn_step(name):
    """get_optimizer(--optimizer sgd|rmsprop|adam) -> CUDA-graph-captured KRN step == the eager step.
    lr is small on purpose: RMSprop / Adam move EVERY parameter by ~lr per step whatever its gradient, and at lr 1e-3 a
    randomly initialised KRN leaves the basin within two steps (loss 12 -> 4e4, measured), where fp32 summation-order noise
    between two runs is amplified to percents and a trajectory comparison says nothing."""
    from types import SimpleNamespace
    from speedplusbaseline_b200.nets.build import get_optimizer
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    cfg = SimpleNamespace(optimizer=name, lr=1e-5, momentum=0.9, weight_decay=0.01, model_name='krn', dann=False)
    x, y = torch.rand(4, 3, 224, 224, generator=_g(1)).cuda(), torch.rand(4, 2, 11, generator=_g(2)).cuda()
    losses, params = [], []
    for use_graph in (True, False):
        m = KeypointRegressionNet(11, device='cuda', seed=3)
        m.train()
        p0 = m._store.params.clone()
        opt = get_optimizer(cfg, m)
        assert type(opt).__name__ == {'sgd': 'FusedSGD', 'rmsprop': 'FusedRMSprop', 'adam': 'FusedAdam'}[name]
        st = KRNTrainStep(m, opt, use_graph=use_graph)
        losses.append([float(st.step(x, y)[0]) for _ in range(3)])
        assert torch.isfinite(m._store.params).all()
        params.append(m._store.params.clone())
        assert float((params[-1] - p0).abs().max()) > 0              # the captured update really ran
    assert all(np.isfinite(losses[0])) and losses[0][2] != losses[0][0]
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-2)      # graph replay == eager (fp32 atomics reorder only)
    assert abs(losses[0][0] - losses[1][0]) <= 1e-4 * abs(losses[1][0])          # first step: same weights, same batch
    assert rel(params[0], params[1]) < 1e-3
