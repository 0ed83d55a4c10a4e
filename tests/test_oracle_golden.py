"""The oracle (oracle/*.py) against outputs of the unmodified reference (tests/golden,
made by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import krn, revgrad, spn, ghiasi, synth, steps

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _close_norms(keys, ref, sd_like, rtol):
    for k, r in zip(keys, ref):
        v = float(sd_like[str(k)].detach().double().norm())
        assert abs(v - r) <= rtol * max(abs(r), 1e-6) + 1e-7, (str(k), v, r)


def test_krn_eval_matches_reference(golden_dir):
    g = _g(golden_dir, 'krn_eval_b2.npz')
    sd = synth.synth_state_dict(krn.krn_shapes(), 2021)
    x = synth.synth_images(2)
    assert synth.checksum(x) == pytest.approx(float(g['x_sum']), rel=1e-12)
    assert synth.checksum(sd['head.0.weight']) == pytest.approx(float(g['w_sum']), rel=1e-12)
    with torch.no_grad():
        xc, yc = krn.krn_forward(sd, x)
    np.testing.assert_allclose(xc.numpy(), g['xc'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(yc.numpy(), g['yc'], rtol=1e-4, atol=1e-5)


def test_krn_train_steps_match_reference(golden_dir):
    g = _g(golden_dir, 'krn_train_b4.npz')
    sd = synth.synth_state_dict(krn.krn_shapes(), 2021)
    st = steps.new_state(sd)
    for i in range(2):
        x, y = synth.synth_images(4, seed=2021 + i), synth.synth_keypoints(4, seed=2021 + i)
        out = steps.krn_train_step(sd, st, x, y)
        np.testing.assert_allclose([out['loss'], out['loss_x'], out['loss_y']], g['losses'][i], rtol=2e-4)
        if i == 0:
            coef = min(1.0, 1.0 / (out['grad_norm'] + 1e-6))
            clipped = {k: v * coef for k, v in out['grads'].items()}
            _close_norms(g['grad_keys'], g['clipped_grad_norms_step1'], clipped, 2e-3)
            _close_norms(g['keys'], g['norms_step1'], sd, 1e-4)
    _close_norms(g['keys'], g['norms_step2'], sd, 1e-4)
    np.testing.assert_allclose(sd['head.0.bias'].numpy(), g['head_bias_step2'], rtol=1e-3, atol=1e-5)
    assert int(sd['base.0.1.num_batches_tracked']) == int(g['nbt'])


def test_dann_epoch_matches_reference(golden_dir):
    g = _g(golden_dir, 'dann_b2.npz')
    sd = synth.synth_state_dict(revgrad.revgrad_shapes(), 2021)
    st = steps.new_state(sd)
    n_b = 3
    for i in range(n_b):
        src, lab = synth.synth_images(2, seed=10 + i), synth.synth_keypoints(2, seed=10 + i)
        tgt = synth.synth_images(2, seed=20 + i, tag='target')
        alpha = revgrad.dann_alpha(i, 1, n_b, 75)
        out = steps.dann_train_step(sd, st, src, lab, tgt, alpha)
    coef = min(1.0, 1.0 / (out['grad_norm'] + 1e-6))
    clipped = {k: v * coef for k, v in out['grads'].items()}
    _close_norms(g['grad_keys'], g['clipped_grad_norms_last'], clipped, 5e-3)
    _close_norms(g['keys'], g['norms'], sd, 1e-4)
    np.testing.assert_allclose(sd['domain_classifier.0.bias'].numpy()[:16], g['dom_bias'], rtol=1e-3, atol=1e-5)
    assert int(sd['net.base.0.1.num_batches_tracked']) == int(g['nbt'])


def test_ghiasi_matches_reference(golden_dir):
    g = _g(golden_dir, 'ghiasi_synth_64.npz')
    sd = synth.synth_state_dict(ghiasi.ghiasi_shapes(), 7)
    x = synth.synth_images(2, 64, 64, seed=7)
    with torch.no_grad():
        out = ghiasi.ghiasi_forward(sd, x, torch.from_numpy(g['style']))
    np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not os.path.exists('/root/reference/src/styleaug/checkpoints/checkpoint_transformer.pth'),
                    reason='real style checkpoint only exists in the build container')
def test_styleaug_real_checkpoint_matches_reference(golden_dir):
    g = _g(golden_dir, 'styleaug_real_64.npz')
    ck = '/root/reference/src/styleaug/checkpoints/'
    sd = torch.load(ck + 'checkpoint_transformer.pth', map_location='cpu', weights_only=False)['state_dict_ghiasi']
    emb = torch.load(ck + 'checkpoint_embeddings.pth', map_location='cpu', weights_only=False)
    base = torch.from_numpy(np.load(ck + 'embedding_mean_speedplus.npy')).float()
    A = ghiasi.style_matrix(emb['pbn_embedding_covariance'].numpy())
    assert synth.checksum(A) == pytest.approx(float(g['A_sum']), rel=1e-6)
    e = ghiasi.mix_embedding(torch.from_numpy(g['noise']), A, emb['pbn_embedding_mean'], base, 0.5)
    x = synth.synth_images(2, 64, 64, seed=7)
    with torch.no_grad():
        out = ghiasi.ghiasi_forward(sd, x, e)
    np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-4, atol=1e-5)


def test_spn_eval_matches_reference(golden_dir):
    g = _g(golden_dir, 'spn_eval_b2.npz')
    sd = synth.synth_state_dict(spn.spn_shapes(), 2021)
    x = synth.synth_images(2, 227, 227)
    with torch.no_grad():
        c, r = spn.spn_forward(sd, x)
    np.testing.assert_allclose(c.numpy(), g['c'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r.numpy(), g['r'], rtol=1e-4, atol=1e-5)
    assert (c.argmax(1).numpy() == g['argmax_c']).all() and (r.argmax(1).numpy() == g['argmax_r']).all()


def test_spn_train_step_matches_reference(golden_dir):
    g = _g(golden_dir, 'spn_train_b2.npz')
    sd = synth.synth_state_dict(spn.spn_shapes(), 2021)
    st = steps.new_state(sd)
    x = synth.synth_images(2, 227, 227)
    out = steps.spn_train_step(sd, st, x, synth.synth_soft_targets(2, tag='cls'),
                               synth.synth_soft_targets(2, tag='wts'), drop_p=0.0)
    clipped = {k: v.clamp(-1, 1) for k, v in out['grads'].items()}
    _close_norms(g['grad_keys'], g['clipped_grad_norms'], clipped, 2e-3)
    _close_norms(g['keys'], g['norms'], sd, 1e-4)
    np.testing.assert_allclose(sd['fc8.bias'].numpy()[:32], g['fc8_bias'], rtol=1e-3, atol=1e-6)
