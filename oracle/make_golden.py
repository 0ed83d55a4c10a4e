"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python -m oracle.make_golden

Shims (no reference file is edited or copied):
  1. sys.argv sanitised before ``import config`` (config.py:64 parses at import)
  2. matplotlib stubbed (src/utils/visualize.py:28-31, absent in this image)
  3. torchvision.models.mobilenet_v2 forced to weights=None (park2019.py:107 downloads)
  4. torch.load forced to map_location='cpu' (styleAugmentor.py:23-24, CUDA storages)
  5. SPN: pretrain=False (bvlc_alexnet.npy absent) and Dropout inplace=False, p=0
Weights/inputs are the seeded synthetic tensors of oracle/synth.py, loaded with
``load_state_dict(strict=True)`` into the reference modules.
"""
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _shims():
    sys.argv = ['make_golden']
    sys.path.insert(0, REF)
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        sys.modules[n] = types.ModuleType(n)
    sys.modules['matplotlib'].use = lambda *a, **k: None
    import torch
    import torchvision.models as tvm
    _mb = tvm.mobilenet_v2
    tvm.mobilenet_v2 = lambda pretrained=False, **kw: _mb(weights=None, **kw)
    _tl = torch.load
    torch.load = lambda f, *a, **k: _tl(f, *a, **{**k, 'map_location': 'cpu', 'weights_only': False})


def _norms(named):
    keys = sorted(named)
    return np.array(keys), np.array([float(named[k].detach().double().norm()) for k in keys])


def main():
    _shims()
    import torch
    from config import cfg
    from src.nets.park2019 import KeypointRegressionNet
    from src.nets.revgrad import RevGrad
    from src.nets.spn import SpacecraftPoseNet
    from src.nets.build import get_optimizer
    from src.core.trainer import train_single_epoch_krn, train_single_epoch_spn
    from src.core.dann import train_dann_single_epoch_krn
    from src.styleaug.ghiasi import Ghiasi
    from src.styleaug.styleAugmentor import StyleAugmentor
    from oracle import krn, revgrad, spn, ghiasi, synth

    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    cfg.optimizer, cfg.lr, cfg.momentum, cfg.weight_decay = 'adamw', 1e-3, 0.9, 0.01
    cfg.use_cuda = False
    dev = torch.device('cpu')

    # ---------------- KRN eval forward (BASELINE config 0) -----------------
    sd = synth.synth_state_dict(krn.krn_shapes(), 2021)
    m = KeypointRegressionNet(11)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = synth.synth_images(2)
    with torch.no_grad():
        xc, yc = m(x)
    np.savez(os.path.join(OUT, 'krn_eval_b2.npz'), xc=xc.numpy(), yc=yc.numpy(),
             x_sum=synth.checksum(x), w_sum=synth.checksum(sd['head.0.weight']))

    # ---------------- KRN two train steps through the reference loop -------
    m = KeypointRegressionNet(11)
    m.load_state_dict(synth.synth_state_dict(krn.krn_shapes(), 2021), strict=True)
    opt = get_optimizer(cfg, m)
    B = 4
    batches = [(synth.synth_images(B, seed=2021 + i), synth.synth_keypoints(B, seed=2021 + i))
               for i in range(2)]
    losses = []
    import copy
    for i, b in enumerate(batches):
        # one-iteration "epoch" so that state can be read after every step
        pre = copy.deepcopy(m.state_dict())
        train_single_epoch_krn(1, cfg, m, [b], opt, None, dev)
        # loss of this step: recompute from the pre-step weights (train-mode BN, same batch)
        m2 = KeypointRegressionNet(11)
        m2.load_state_dict(pre)
        m2.train()
        l, sm = m2(b[0], b[1])
        losses.append([float(l), sm['loss_x'], sm['loss_y']])
        if i == 0:
            gk, gn = _norms({k: p.grad for k, p in m.named_parameters()})
            pk1, pn1 = _norms(m.state_dict())
    pk, pn = _norms(m.state_dict())
    np.savez(os.path.join(OUT, 'krn_train_b4.npz'), losses=np.array(losses),
             grad_keys=gk, clipped_grad_norms_step1=gn, keys=pk, norms_step1=pn1, norms_step2=pn,
             head_bias_step2=m.state_dict()['head.0.bias'].numpy(),
             nbt=int(m.state_dict()['base.0.1.num_batches_tracked']))

    # ---------------- DANN step through the reference loop -----------------
    cfg.max_epochs = 75
    r = RevGrad(11)
    r.load_state_dict(synth.synth_state_dict(revgrad.revgrad_shapes(), 2021), strict=True)
    opt = get_optimizer(cfg, r)
    B = 2
    n_b = 3
    src = [(synth.synth_images(B, seed=10 + i), synth.synth_keypoints(B, seed=10 + i)) for i in range(n_b)]
    tgt = [synth.synth_images(B, seed=20 + i, tag='target') for i in range(n_b)]
    train_dann_single_epoch_krn(1, cfg, r, src, tgt, opt, None, dev)
    gk, gn = _norms({k: p.grad for k, p in r.named_parameters()})
    pk, pn = _norms(r.state_dict())
    np.savez(os.path.join(OUT, 'dann_b2.npz'), grad_keys=gk, clipped_grad_norms_last=gn, keys=pk, norms=pn,
             dom_bias=r.state_dict()['domain_classifier.0.bias'].numpy()[:16],
             nbt=int(r.state_dict()['net.base.0.1.num_batches_tracked']))

    # ---------------- Ghiasi / StyleAugmentor ------------------------------
    g = Ghiasi()
    gsd = synth.synth_state_dict(ghiasi.ghiasi_shapes(), 7)
    g.load_state_dict(gsd, strict=True)
    g.eval()
    x = synth.synth_images(2, 64, 64, seed=7)
    style = torch.randn(2, 100, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        out = g(x, style)
    np.savez(os.path.join(OUT, 'ghiasi_synth_64.npz'), out=out.numpy(), style=style.numpy())
    # real checkpoint, through StyleAugmentor (styleAugmentor.py:44-68)
    aug = StyleAugmentor(0.5, dev)
    torch.manual_seed(123)
    noise = torch.randn(2, 100)
    torch.manual_seed(123)
    out = aug(x)
    np.savez(os.path.join(OUT, 'styleaug_real_64.npz'), out=out.numpy(), noise=noise.numpy(),
             A_sum=synth.checksum(aug.A))

    # ---------------- SPN ---------------------------------------------------
    ssd = synth.synth_state_dict(spn.spn_shapes(), 2021)
    s = SpacecraftPoseNet(5000, pretrain=False)
    s.load_state_dict(ssd, strict=True)
    s.eval()
    x = synth.synth_images(2, 227, 227)
    with torch.no_grad():
        c, rr = s(x)
    np.savez(os.path.join(OUT, 'spn_eval_b2.npz'), c=c.numpy(), r=rr.numpy(),
             argmax_c=c.argmax(1).numpy(), argmax_r=rr.argmax(1).numpy())
    for n in (6, 7, 9, 10):
        d = getattr(s, 'dropout%d' % n)
        d.inplace, d.p = False, 0.0
    opt = get_optimizer(cfg, s)
    yc_, yw_ = synth.synth_soft_targets(2, tag='cls'), synth.synth_soft_targets(2, tag='wts')
    train_single_epoch_spn(1, cfg, s, [(x, yc_, yw_)], opt, None, dev)
    gk, gn = _norms({k: p.grad for k, p in s.named_parameters()})
    pk, pn = _norms(s.state_dict())
    np.savez(os.path.join(OUT, 'spn_train_b2.npz'), grad_keys=gk, clipped_grad_norms=gn, keys=pk, norms=pn,
             fc8_bias=s.state_dict()['fc8.bias'].numpy()[:32])
    print('golden written to', OUT)


if __name__ == '__main__':
    main()
