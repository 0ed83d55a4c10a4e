"""Shared pieces of the train.py / adapt.py / test.py front-ends (reference train.py:49-158,
adapt.py:47-146, test.py:42-88): device selection, the --no_cuda escape to the reference's own torch
modules, loaders (SPEED+ through the reference's src.datasets when a checkout is given, synthetic
otherwise -- datasets are out of scope), checkpoint resume."""
import logging
import os
import sys

import torch

logger = logging.getLogger(__name__)


def select_device(cfg):
    """cuda:0 like the reference (train.py:50); under torchrun one process per GPU on cuda:LOCAL_RANK with an NCCL group."""
    from .dist import setup_cli
    return setup_cli(cfg)


def setup_logger(name):
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(name)s %(levelname)s %(message)s')


def reference_modules(cfg):
    """import the reference checkout (for --no_cuda, datasets and evaluation)."""
    root = cfg.reference_root
    if not root or not os.path.isdir(root):
        raise SystemExit('this option needs --reference_root <checkout of tpark94/speedplusbaseline> '
                         '(datasets / evaluation / the --no_cuda torch path live there)')
    if root not in sys.path:
        sys.path.insert(1, root)
    return root


class SyntheticLoader:
    """`n` batches of seeded synthetic data shaped like the SPEED+ loaders' output
    (Park2019KRNDataset.py:81-109: image [B,3,H,W] in [0,1], keypoints [B,2,K])."""

    def __init__(self, cfg, n, labels=True, seed=0, pin=True):
        self.n, self.labels = n, labels
        g = torch.Generator().manual_seed(cfg.seed + seed + 1000 * getattr(cfg, 'rank', 0))     # every rank its own shard
        H, W = cfg.input_shape[0], cfg.input_shape[1]
        self.images = torch.rand(cfg.batch_size, 3, H, W, generator=g)
        self.target = torch.rand(cfg.batch_size, 2, cfg.num_keypoints, generator=g)
        self.spn = cfg.model_name == 'spn' and not cfg.dann
        if self.spn:        # SPNDataset.py:83-94: soft n-hot attitude-class / weight targets
            t = torch.zeros(cfg.batch_size, cfg.num_classes)
            for b in range(cfg.batch_size):
                idx = torch.randperm(cfg.num_classes, generator=g)[:cfg.num_neighbors]
                t[b, idx] = 1.0 / cfg.num_neighbors
            self.yc, self.yw = t, t.clone()
        if pin and torch.cuda.is_available():
            self.images, self.target = self.images.pin_memory(), self.target.pin_memory()

    def __len__(self):
        return self.n

    def __iter__(self):
        for _ in range(self.n):
            if self.spn:
                yield (self.images, self.yc, self.yw)
            else:
                yield (self.images, self.target) if self.labels else self.images


def make_loaders(cfg, specs):
    """specs: list of dict(is_train, is_source, load_labels) as passed to the reference's make_dataloader
    (src/datasets/build.py:45-66)."""
    if cfg.synthetic_data > 0:
        return [SyntheticLoader(cfg, cfg.synthetic_data, labels=s.get('load_labels', True), seed=i) for i, s in enumerate(specs)]
    on_device = getattr(cfg, 'device_transforms', False) and torch.cuda.is_available() and cfg.use_cuda
    out = []
    for s in specs:
        # SPN TRAINING batches carry soft attitude-class targets built by the reference's SPNDataset (SPNDataset.py:83-94):
        # those loaders stay with the reference; everything else (KRN train / DANN target / test, SPN test) can decode-only
        if on_device and not (cfg.model_name == 'spn' and s.get('is_train', True)):
            from .datasets.raw import make_dataloader as make_device_dataloader
            out.append(make_device_dataloader(cfg, device=getattr(cfg, 'device', None) or torch.device('cuda:0'), **s))
        else:
            reference_modules(cfg)
            from src.datasets.build import make_dataloader
            out.append(make_dataloader(cfg, **s))
    return out


def resume(cfg, model, optimizer, device):
    from .utils import load_checkpoint
    f = os.path.join(cfg.savedir, 'checkpoint.pth.tar')
    if cfg.auto_resume and os.path.exists(f):
        last_epoch, _ = load_checkpoint(f, model, optimizer, device)
        return last_epoch
    return 0


def validate(cfg, model, test_loader, epoch, writer, device):
    """Evaluation (src/core/inference.py:43-221).  On the GPU: batched forward + on-device post-processing tail
    (core/inference.py); with --no_cuda: the reference's own loop.  EPnP / SPEED metrics are the
    reference's CPU code either way (out of scope)."""
    reference_modules(cfg)
    from scipy.io import loadmat
    from src.utils.utils import load_tango_3d_keypoints, load_camera_intrinsics
    corners3D = load_tango_3d_keypoints(os.path.join(cfg.reference_root, cfg.keypts_3d_model))
    cameraMatrix, distCoeffs = load_camera_intrinsics(os.path.join(cfg.dataroot, cfg.dataname, 'camera.json'))
    att = loadmat(os.path.join(cfg.reference_root, cfg.attitude_class))['qClass']
    if device.type == 'cuda':
        from .core import inference
    else:
        from src.core import inference
    fn = getattr(inference, 'valid_' + cfg.model_name)
    return fn(epoch, cfg, model, test_loader, cameraMatrix, distCoeffs, corners3D, writer, device, att)
