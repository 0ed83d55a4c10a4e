#!/usr/bin/env python
"""test.py -- evaluation CLI (reference test.py:42-88): loads the bare state_dict written to
model_best.pth.tar (utils.py:114-118) into the B200 model and runs the reference's validation loop
(EPnP / SPEED metrics are CPU post-processing and out of scope: they come from --reference_root).
With --synthetic_data N it instead times N batch forwards and prints the keypoint logits' checksum."""
import logging
import os

import torch

from config import cfg

logger = logging.getLogger(__name__)


def main():
    from speedplusbaseline_b200 import cli
    from speedplusbaseline_b200.utils import set_all_seeds
    device = cli.select_device(cfg)
    cli.setup_logger('test')
    set_all_seeds(cfg.seed, cfg, True)
    if device.type == 'cuda':
        from speedplusbaseline_b200.nets.build import get_model
    else:
        cli.reference_modules(cfg)
        from src.nets.build import get_model
    model = get_model(cfg)
    if cfg.pretrained:
        sd = torch.load(cfg.pretrained, map_location='cpu', weights_only=False)
        model.load_state_dict(sd, strict=True)
        logger.info('Loaded {}'.format(cfg.pretrained))
    model.to(device)
    model.eval()
    (test_loader,) = cli.make_loaders(cfg, [dict(is_train=False, is_source=False, load_labels=True)])
    if cfg.synthetic_data > 0:
        import time
        t0, s = time.time(), 0.0
        with torch.no_grad():
            for images, _ in test_loader:
                out = model(images.to(device))
                xc = out[0] if isinstance(out, tuple) else out
                s += float(xc.double().sum())
        dt = time.time() - t0
        print('forward of %d batches of %d: %.1f images/s, keypoint-x checksum %.6f'
              % (len(test_loader), cfg.batch_size, len(test_loader) * cfg.batch_size / dt, s))
        return
    cli.validate(cfg, model, test_loader, 0, None, device)


if __name__ == '__main__':
    main()
