"""SURVEY.md 8(f) rows on the CPU: the oracle restatements of the non-AdamW optimizers against fixtures produced by
the reference's own get_optimizer (oracle/make_golden_optim.py -> tests/golden/optim_steps.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import optim as ooptim
from oracle.make_golden_optim import HYPER, STEPS, synth_problem


def _run_oracle(name, dtype=torch.float32):
    p0, grads = synth_problem()
    params = [p.clone().to(dtype) for p in p0]
    st = ooptim.AdamWState(params) if name == 'adamw' else ooptim.FlatState(params)
    traj = []
    for s in range(STEPS):
        gl = [g.clone().to(dtype) for g in grads[s]]
        ooptim.clip_grad_norm(gl, 1.0)
        if name == 'adamw':
            ooptim.adamw_step(params, gl, st, lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, wd=HYPER['weight_decay'])
        elif name == 'adam':
            ooptim.adam_step(params, gl, st, lr=HYPER['lr'], beta1=HYPER['momentum'], beta2=0.999, eps=1e-8, wd=HYPER['weight_decay'])
        elif name == 'sgd':
            ooptim.sgd_step(params, gl, st, lr=HYPER['lr'], momentum=HYPER['momentum'], wd=HYPER['weight_decay'])
        else:
            ooptim.rmsprop_step(params, gl, st, lr=HYPER['lr'], alpha=HYPER['momentum'], eps=1e-8, wd=HYPER['weight_decay'])
        traj.append(torch.cat([p.reshape(-1) for p in params]).double().numpy().copy())
    return np.stack(traj)


@pytest.mark.parametrize('name', ['sgd', 'rmsprop', 'adam', 'adamw'])
def test_oracle_optimizers_match_reference_factory(golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'optim_steps.npz'))[name]
    got = _run_oracle(name)
    # same fp32 arithmetic as torch.optim: agreement to a few ulps of the parameter scale
    np.testing.assert_allclose(got, g, rtol=2e-6, atol=2e-7)
    # the trajectories are not trivially equal to the start (every step moved the parameters)
    p0 = torch.cat([p.reshape(-1) for p in synth_problem()[0]]).numpy()
    assert np.abs(g[0] - p0).max() > 1e-4 and np.abs(g[-1] - g[0]).max() > 1e-4


# ---- row f2: evaluation tail ----------------------------------------------------------------------
def _tie_free_rows(w, k):
    """rows whose k+1 largest values are pairwise distinct (torch.topk's order among exact ties is unspecified)."""
    s = -np.sort(-w, axis=1)[:, :k + 1]
    return np.array([len(np.unique(r)) == k + 1 for r in s])


def test_oracle_postproc_matches_reference_loops(golden_dir):
    from oracle import postproc
    from oracle.make_golden_postproc import K_NB, synth_inputs
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    w, x, y, bb = (t.numpy() for t in synth_inputs())
    tw, ti = postproc.spn_top_classes(w, K_NB)
    free = _tie_free_rows(w, K_NB)
    assert free.sum() >= 4 and not free.all()              # the fixture holds both kinds of rows
    assert (ti[free] == g['top_idx'][free]).all()          # bit-exact indices where the order is defined
    # every row, ties included: the selected VALUES are the reference's, in the same (descending) order
    assert (np.take_along_axis(w, ti, 1) == np.take_along_axis(w, g['top_idx'], 1)).all()
    np.testing.assert_allclose(tw, g['top_w'], rtol=1e-6, atol=1e-9)
    assert (postproc.krn_keypoints_pix(x, y, bb) == g['kpt_pix']).all()       # fp32, bit-exact


# ---- row f3: pretrained-weight importers ------------------------------------------------------------
def _tv_mobilenet_sd(seed=3):
    import torchvision
    torch.manual_seed(seed)
    m = torchvision.models.mobilenet_v2(weights=None)
    for mod in m.modules():                       # non-trivial BN statistics, like a trained checkpoint
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
            mod.num_batches_tracked.fill_(7)
    return m.state_dict()


@pytest.mark.parametrize('dann', [False, True])
def test_mobilenetv2_importer_fills_every_backbone_key(dann):
    from speedplusbaseline_b200 import importers
    from speedplusbaseline_b200.krn_engine import krn_layout
    from speedplusbaseline_b200.params import ParamStore
    prefix = 'net.' if dann else ''
    W, BN, order = krn_layout(11, prefix=prefix, dann=dann)
    st = ParamStore(W, BN, torch.device('cpu'))
    before = st.state_dict(order)
    tv = _tv_mobilenet_sd()
    written = importers.load_mobilenetv2_backbone(st, tv)
    after = st.state_dict(order)
    base_keys = [k for k in order if k.startswith(prefix + 'base.')]
    assert sorted(written) == sorted(base_keys)                         # every backbone key, nothing else
    for k in order:
        if k in written:
            src = tv['features.' + k[len(prefix) + len('base.'):]]
            assert torch.equal(after[k], src.to(after[k].dtype)), k    # values survive the native-layout round trip
        else:
            assert torch.equal(after[k], before[k]), k                 # extras / head / domain classifier untouched
    assert not any(k.startswith(prefix + 'base.18') for k in written)
    with pytest.raises(KeyError):
        importers.mobilenetv2_to_krn({k: v for k, v in tv.items() if not k.startswith('features.7.')})


def test_alexnet_npy_importer_matches_reference_transposition(tmp_path):
    from speedplusbaseline_b200 import importers
    from speedplusbaseline_b200.params import ParamStore
    from speedplusbaseline_b200.spn_engine import CONVS, spn_layout
    rng = np.random.default_rng(5)
    dump = {}
    for name, ci, co, k, s, p, g in CONVS:       # Caffe layout [H, W, Cin/groups, Cout] + bias, bytes keys like the real file
        dump[name.encode()] = [rng.standard_normal((k, k, ci // g, co)).astype(np.float32), rng.standard_normal(co).astype(np.float32)]
    dump[b'fc6'] = [rng.standard_normal((9216, 16)).astype(np.float32), rng.standard_normal(16).astype(np.float32)]   # ignored
    path = os.path.join(str(tmp_path), 'bvlc_alexnet.npy')
    np.save(path, dump, allow_pickle=True)
    W, order = spn_layout(40)
    st = ParamStore(W, [], torch.device('cpu'))
    written = importers.load_alexnet_npy(st, path)
    assert sorted(written) == sorted(n + s for n, *_ in CONVS for s in ('.weight', '.bias'))
    sd = st.state_dict(order)
    for name, *_ in CONVS:
        w, b = dump[name.encode()]
        # spn.py:118-121: np.transpose(data, (3, 2, 0, 1)) -> conv.weight
        assert torch.equal(sd[name + '.weight'], torch.from_numpy(np.transpose(w, (3, 2, 0, 1)).copy()))
        assert torch.equal(sd[name + '.bias'], torch.from_numpy(b))
    assert float(sd['fc6.weight'].abs().sum()) == 0.0                    # FC layers are not part of the dump the reference reads


@pytest.mark.skipif(not os.path.isdir('/root/reference/src'), reason='needs the reference checkout (build container only)')
def test_mobilenetv2_importer_agrees_with_reference_module():
    """the reference builds `base` from torchvision's features[:-1] (park2019.py:107-108): with the same RNG stream its
    state_dict must equal the importer's mapping of the torchvision checkpoint, key for key."""
    import subprocess
    import sys
    code = r'''
import sys, types
sys.argv = ['x']; sys.path.insert(0, '/root/reference'); sys.path.insert(0, %r)
for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
    sys.modules[n] = types.ModuleType(n)
sys.modules['matplotlib'].use = lambda *a, **k: None
import torch, torchvision.models as tvm
_mb = tvm.mobilenet_v2
tvm.mobilenet_v2 = lambda pretrained=False, **kw: _mb(weights=None, **kw)
from src.nets.park2019 import KeypointRegressionNet
from speedplusbaseline_b200 import importers
torch.manual_seed(11); ref = KeypointRegressionNet(11).state_dict()
torch.manual_seed(11); tv = _mb(weights=None).state_dict()
mapped = importers.mobilenetv2_to_krn(tv)
assert sorted(mapped) == sorted(k for k in ref if k.startswith('base.')), 'key sets differ'
assert all(torch.equal(mapped[k], ref[k]) for k in mapped)
print('OK', len(mapped))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, B200SP_NO_AUTOBUILD='1'))
    assert r.returncode == 0 and 'OK 306' in r.stdout, r.stdout + r.stderr


# ---- row f1: input pipeline (oracle + host logic) ---------------------------------------------------
def test_resampler_restatement_is_bit_exact_vs_pillow():
    """third-party arithmetic (Pillow's 8-bit BILINEAR resize, used through T.resized_crop): the restatement against
    the library itself, up- and down-scaling, one-axis-only, identity, RGB and grey."""
    from PIL import Image
    from oracle import transforms as ot
    rng = np.random.default_rng(0)
    for H, W, oh, ow, C in [(300, 300, 224, 224, 3), (1200, 1100, 224, 224, 1), (97, 101, 224, 224, 3), (224, 500, 224, 224, 1),
                            (500, 224, 224, 224, 3), (224, 224, 224, 224, 3), (640, 481, 227, 227, 1), (50, 1000, 224, 224, 3)]:
        img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
        pil = Image.fromarray(img if C == 3 else img[:, :, 0]).resize((ow, oh), Image.BILINEAR)
        ref = np.asarray(pil).reshape(oh, ow, C)
        assert np.array_equal(ot.pil_resize_bilinear_u8(img, oh, ow), ref), (H, W, oh, ow, C)


def _replay_case(case):
    from oracle import transforms as ot
    from oracle.make_golden_transforms import FRAME_HW, synth_frame, synth_keypoints
    seed, bbox, p, is_train, model = case
    H, W = FRAME_HW
    frame = np.repeat(synth_frame(seed)[:, :, None], 3, 2)
    kp = synth_keypoints(seed, bbox)
    gen = torch.Generator().manual_seed(seed)
    size = (224, 224) if model == 'krn' else (227, 227)
    dec = dict(rot=0, flip=0, bc=None, noise=None)
    if model == 'krn':
        u = [torch.rand(1, generator=gen) for _ in range(3)] if is_train else None
        box = ot.random_crop_box(bbox, W, H, is_train, u)
        k, bb = ot.crop_keypoints(kp, box), np.array(box, np.float32)
    else:
        box = ot.resize_crop_box(bbox, W, H)
        k, bb = torch.as_tensor(kp), np.array(bbox, np.float32)
    img = ot.crop_resize_to_tensor(frame, box, size)
    if is_train and model == 'krn':
        dec = ot.draw_augment(p, img.shape, gen)
        img, k = ot.apply_augment(img, k, **dec)
    return img, bb, k, box, dec


def test_transforms_oracle_replays_reference_bit_exact(golden_dir):
    from oracle.make_golden_transforms import CASES
    g = np.load(os.path.join(golden_dir, 'transforms.npz'))
    fired = set()
    for case in CASES:
        img, bb, k, box, dec = _replay_case(case)
        s = case[0]
        assert np.array_equal(img.numpy(), g['img%d' % s]), s
        assert np.array_equal(bb, g['bbox%d' % s]) and np.array_equal(k.numpy(), g['kpt%d' % s]), s
        fired |= {n for n in ('rot', 'flip') if dec[n]} | {n for n in ('bc', 'noise') if dec[n] is not None}
    assert fired == {'rot', 'flip', 'bc', 'noise'}          # the fixture exercises every augmentation


def test_host_sampler_follows_reference_draws():
    """datasets/transforms.py (product host code) draws the same decisions as the oracle replay of the reference stream."""
    from oracle import transforms as ot
    from oracle.make_golden_transforms import CASES, FRAME_HW
    from speedplusbaseline_b200.datasets import transforms as dt
    H, W = FRAME_HW
    for seed, bbox, p, is_train, model in CASES:
        if model != 'krn':
            continue
        g1, g2 = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(seed)
        box = dt.sample_crop_box(bbox, W, H, is_train, g1)
        u = [torch.rand(1, generator=g2) for _ in range(3)] if is_train else None
        assert box == ot.random_crop_box(bbox, W, H, is_train, u)
        if is_train:
            rot, flip, bc, a, b, std = dt.sample_augment(p, g1)
            d = ot.draw_augment(p, (3, 224, 224), g2)
            assert (rot, flip, bool(bc), std > 0) == (d['rot'], d['flip'], d['bc'] is not None, d['noise'] is not None)
            if bc:
                assert a == float(d['bc'][0]) and b == float(d['bc'][1])


def test_dest_index_matches_torch_rot90_and_flip():
    from speedplusbaseline_b200.datasets.transforms import dest_index
    n = 5
    x = torch.arange(n * n).reshape(1, n, n)
    for rot in range(4):
        for flip in range(3):
            ref = torch.rot90(x, rot, (1, 2)) if rot else x
            ref = ref.flip(2) if flip == 1 else (ref.flip(1) if flip == 2 else ref)
            out = torch.empty_like(x)
            for sy in range(n):
                for sx in range(n):
                    i, j = dest_index(sy, sx, n, n, rot, flip)
                    out[0, i, j] = x[0, sy, sx]
            assert torch.equal(out, ref), (rot, flip)


def test_vectorised_sampler_equals_per_sample_formulas():
    """sample_batch (one block of uniforms per batch) == the per-sample reference-order functions fed the same uniforms."""
    from oracle import transforms as ot
    from speedplusbaseline_b200.datasets import transforms as dt
    g = torch.Generator().manual_seed(5)
    B, W, H = 64, 1920, 1200
    c = torch.rand(B, 2, generator=g) * torch.tensor([1920., 1200.])
    sz = torch.rand(B, 2, generator=g) * 900 + 20
    bbox = torch.stack([c[:, 0] - sz[:, 0] / 2, c[:, 0] + sz[:, 0] / 2, c[:, 1] - sz[:, 1] / 2, c[:, 1] + sz[:, 1] / 2], 1).numpy()
    u = torch.rand(B, 11, generator=g)
    for is_train in (True, False):
        d = dt.sample_batch(bbox, W, H, 'krn', is_train, 0.5, u=u)
        for i in range(B):
            assert d['box'][i] == ot.random_crop_box(bbox[i], W, H, is_train, [u[i, 0:1], u[i, 1:2], u[i, 2:3]]), i
        if is_train:
            la, lb = torch.tensor((0.5, 2.0)).log(), torch.tensor((-25, 25)) / 255
            for i in range(B):
                assert d['rot'][i] == (1 + min(2, int(u[i, 4] * 3)) if u[i, 3] < 0.5 else 0)
                assert d['flip'][i] == ((1 if u[i, 6] < 0.5 else 2) if u[i, 5] < 0.5 else 0)
                if u[i, 7] < 0.5:
                    assert d['bc'][i] == 1
                    assert np.float32(d['a'][i]) == np.float32(float((u[i, 8:9] * (la[1] - la[0]) + la[0]).exp()))
                    assert np.float32(d['b'][i]) == np.float32(float(u[i, 9:10] * (lb[1] - lb[0]) + lb[0]))
                else:
                    assert (d['bc'][i], d['a'][i], d['b'][i]) == (0, 1.0, 0.0)
                assert (d['std'][i] > 0) == bool(u[i, 10] < 0.5)
            assert {0, 1, 2, 3} == set(d['rot']) and {0, 1, 2} == set(d['flip'])
        else:
            assert set(d['rot']) == {0} and set(d['std']) == {0.0}
    s = dt.sample_batch(bbox, W, H, 'spn', False, 0.0, u=u)
    for i in range(B):
        assert s['box'][i] == ot.resize_crop_box(bbox[i], W, H)


def _write_split(root, n, hw=(120, 160), K=11, rgb_index=None):
    """a miniature SPEED+-style tree: <root>/speedplus/synthetic/{images,splits_krn/train.csv} (+ lightbox for tests)."""
    from PIL import Image
    import pandas as pd
    rng = np.random.default_rng(0)
    rows = []
    for dom, csv in (('synthetic', 'train.csv'), ('lightbox', 'lightbox.csv')):
        os.makedirs(os.path.join(root, 'speedplus', dom, 'images'), exist_ok=True)
        os.makedirs(os.path.join(root, 'speedplus', dom, 'splits_krn'), exist_ok=True)
        rows = []
        for i in range(n):
            arr = rng.integers(0, 256, hw, dtype=np.uint8)
            rel = os.path.join(dom, 'images', 'img%03d.png' % i)
            im = Image.fromarray(arr)
            if rgb_index == i:
                im = im.convert('RGB')
            im.save(os.path.join(root, 'speedplus', rel))
            bbox = [20 + i, 100 + i, 10 + i, 90 + i]
            pose = list(rng.normal(size=7))
            kp = list(rng.uniform(20, 100, 2 * K))
            rows.append([rel] + bbox + pose + kp)
        pd.DataFrame(rows).to_csv(os.path.join(root, 'speedplus', dom, 'splits_krn', csv), header=False, index=False)
    return rows


def test_raw_frame_dataset_reads_reference_csv_layout(tmp_path):
    from types import SimpleNamespace
    from PIL import Image
    from speedplusbaseline_b200.datasets.raw import RawFrameDataset
    rows = _write_split(str(tmp_path), 3, rgb_index=1)
    cfg = SimpleNamespace(dataroot=str(tmp_path), dataname='speedplus', num_keypoints=11, model_name='krn', train_domain='synthetic',
                          test_domain='lightbox', train_csv='train.csv', test_csv='lightbox.csv')
    tr = RawFrameDataset(cfg, is_train=True, is_source=True, load_labels=True)
    assert len(tr) == 3
    frame, bbox, kp, q, t = tr[2]
    assert frame.dtype == torch.uint8 and tuple(frame.shape) == (120, 160) and tuple(kp.shape) == (2, 11)
    ref = np.array(Image.open(os.path.join(str(tmp_path), 'speedplus', 'synthetic', 'images', 'img002.png')))
    assert np.array_equal(frame.numpy(), ref)
    assert tuple(tr[1][0].shape) == (120, 160, 3)                               # non-grey files go through RGB like upstream
    te = RawFrameDataset(cfg, is_train=False, is_source=False, load_labels=True)
    f2, b2, k2, q2, t2 = te[0]
    assert np.allclose(b2.numpy(), rows[0][1:5]) and np.allclose(q2.numpy(), rows[0][5:9], atol=1e-6)
    assert np.allclose(t2.numpy(), rows[0][9:12], atol=1e-6) and float(k2.abs().sum()) == 0.0
    # keypoint columns: kx1, ky1, kx2, ... -> [2, K] (Park2019KRNDataset.py:92-93)
    tgt = RawFrameDataset(cfg, is_train=True, is_source=False, load_labels=False)
    assert len(tgt) == 3
    with pytest.raises(AssertionError):
        RawFrameDataset(cfg, is_train=True, is_source=False, load_labels=True)
