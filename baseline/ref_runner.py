"""Run the UNMODIFIED reference (staged at baseline/_ref by tools/stage_reference.py) through its own public
API -- get_model / get_optimizer / train_single_epoch_krn / train_dann_single_epoch_krn / StyleAugmentor -- on the
host CPU (`--no_cuda` semantics) or on cuda:0 (the reference's stock cuDNN path: the "existing Blackwell kernel" bar of
SURVEY.md section 8d).  Nothing of this repo's product code is imported here.

Environment shims (no reference file is edited):
  1. sys.argv is replaced while `config` is imported (config.py:64 parses argv at import);
  2. matplotlib is stubbed (src/utils/visualize.py:28-31 imports it; absent in this image);
  3. torchvision.models.mobilenet_v2 is forced to weights=None (park2019.py:107 would download ImageNet weights);
  4. torch.load gets weights_only=False (+ map_location='cpu' on a CUDA-less host: the style checkpoint holds CUDA storages).
The data loader is the list `[(images, target)] * steps` of synthetic tensors (the reference loops iterate any iterable
with __len__); report_progress prints are silenced.
"""
import contextlib
import io
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF, 'src', 'core', 'trainer.py'))


_state = {}


def _import_reference(argv=()):
    if _state:
        return _state
    if not available():
        raise RuntimeError('baseline/_ref is not staged: run `python tools/stage_reference.py` in the build container')
    import torch
    for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules['matplotlib'], 'use'):
        sys.modules['matplotlib'].use = lambda *a, **k: None
    import torchvision.models as tvm
    _mb = tvm.mobilenet_v2
    tvm.mobilenet_v2 = lambda pretrained=False, **kw: _mb(weights=None, **kw)
    _tl = torch.load

    def _load(f, *a, **k):
        k.setdefault('weights_only', False)
        if not torch.cuda.is_available():
            k.setdefault('map_location', 'cpu')
        return _tl(f, *a, **k)
    torch.load = _load
    sys.path.insert(0, REF)
    old = sys.argv
    sys.argv = ['reference'] + list(argv)
    try:
        from config import cfg
    finally:
        sys.argv = old
    from src.nets.build import get_model, get_optimizer
    from src.core.trainer import train_single_epoch_krn
    from src.core.dann import train_dann_single_epoch_krn
    from src.styleaug.styleAugmentor import StyleAugmentor
    _state.update(cfg=cfg, get_model=get_model, get_optimizer=get_optimizer, train_krn=train_single_epoch_krn,
                  train_dann=train_dann_single_epoch_krn, StyleAugmentor=StyleAugmentor)
    return _state


class _Loader(list):
    """`[(images, target)] * n` with the len() the reference loops ask for."""


def _sync(device):
    import torch
    if device.type == 'cuda':
        torch.cuda.synchronize(device)


def time_krn_train(device='cpu', batch=48, hw=224, warmup=2, steps=5, fp16=False, style=False, threads=None, seed=2021):
    """Times `steps` iterations of the reference's own train_single_epoch_krn (trainer.py:41-112) after `warmup`
    iterations of the same loop.  Returns dict(ms_per_step, images_per_sec, warmup, steps, ...) -- exactly what ran."""
    import torch
    R = _import_reference()
    cfg = R['cfg']
    dev = torch.device(device)
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    cfg.model_name, cfg.dann, cfg.num_keypoints = 'krn', False, 11
    cfg.optimizer, cfg.lr, cfg.weight_decay, cfg.momentum = 'adamw', 1e-3, 0.01, 0.9
    cfg.use_cuda = dev.type == 'cuda'
    cfg.texture_ratio = 1.0
    model = R['get_model'](cfg).to(dev)
    opt = R['get_optimizer'](cfg, model)
    scaler = torch.cuda.amp.GradScaler() if (fp16 and dev.type == 'cuda') else None
    aug = R['StyleAugmentor'](0.5, dev) if style else None
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, hw, hw, generator=g)
    y = torch.rand(batch, 2, 11, generator=g)
    if dev.type == 'cuda':
        x, y = x.pin_memory(), y.pin_memory()
    sink = io.StringIO()

    def run(n):
        with contextlib.redirect_stdout(sink):
            R['train_krn'](1, cfg, model, _Loader([(x, y)] * n), opt, None, dev, styleAugmentor=aug, scaler=scaler)
    if warmup > 0:
        run(warmup)
    _sync(dev)
    t0 = time.perf_counter()
    run(steps)
    _sync(dev)
    dt = (time.perf_counter() - t0) / steps
    return {'ms_per_step': dt * 1e3, 'images_per_sec': batch / dt, 'warmup': warmup, 'steps': steps, 'batch': batch,
            'device': str(dev), 'fp16_autocast': bool(scaler is not None), 'style_aug': bool(style),
            'threads': torch.get_num_threads(), 'torch': torch.__version__,
            'cudnn': torch.backends.cudnn.version() if dev.type == 'cuda' else None,
            'api': 'src.core.trainer.train_single_epoch_krn (unmodified, baseline/_ref)'}


def time_dann_train(device='cpu', batch=48, hw=224, warmup=2, steps=5, threads=None, seed=2021):
    """Same for adapt.py's loop: train_dann_single_epoch_krn (dann.py:38-117), `batch` source + `batch` target images."""
    import torch
    R = _import_reference()
    cfg = R['cfg']
    dev = torch.device(device)
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    cfg.model_name, cfg.dann, cfg.num_keypoints = 'krn', True, 11
    cfg.optimizer, cfg.lr, cfg.weight_decay, cfg.momentum = 'adamw', 1e-3, 0.01, 0.9
    cfg.use_cuda = dev.type == 'cuda'
    cfg.max_epochs = 75
    model = R['get_model'](cfg).to(dev)
    opt = R['get_optimizer'](cfg, model)
    g = torch.Generator().manual_seed(seed)
    xs, xt = torch.rand(batch, 3, hw, hw, generator=g), torch.rand(batch, 3, hw, hw, generator=g)
    y = torch.rand(batch, 2, 11, generator=g)
    if dev.type == 'cuda':
        xs, xt, y = xs.pin_memory(), xt.pin_memory(), y.pin_memory()
    sink = io.StringIO()

    def run(n):
        with contextlib.redirect_stdout(sink):
            R['train_dann'](1, cfg, model, _Loader([(xs, y)] * n), _Loader([xt] * n), opt, None, dev)
    if warmup > 0:
        run(warmup)
    _sync(dev)
    t0 = time.perf_counter()
    run(steps)
    _sync(dev)
    dt = (time.perf_counter() - t0) / steps
    return {'ms_per_step': dt * 1e3, 'images_per_sec': 2 * batch / dt, 'source_images_per_sec': batch / dt,
            'warmup': warmup, 'steps': steps, 'batch': batch, 'device': str(dev), 'threads': torch.get_num_threads(),
            'api': 'src.core.dann.train_dann_single_epoch_krn (unmodified, baseline/_ref)'}


def time_styleaug(device='cuda:0', batch=48, hw=224, warmup=2, steps=5, seed=2021):
    import torch
    R = _import_reference()
    dev = torch.device(device)
    aug = R['StyleAugmentor'](0.5, dev)
    x = torch.rand(batch, 3, hw, hw, generator=torch.Generator().manual_seed(seed)).to(dev)
    for _ in range(warmup):
        aug(x)
    _sync(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        aug(x)
    _sync(dev)
    dt = (time.perf_counter() - t0) / steps
    return {'ms': dt * 1e3, 'images_per_sec': batch / dt, 'warmup': warmup, 'steps': steps, 'device': str(dev),
            'api': 'src.styleaug.styleAugmentor.StyleAugmentor.forward (unmodified, real checkpoints)'}


if __name__ == '__main__':
    import json
    dev = sys.argv[1] if len(sys.argv) > 1 else 'cpu'
    print(json.dumps(time_krn_train(dev, batch=int(os.environ.get('B', '4')), warmup=1, steps=2)))
