#!/usr/bin/env python
"""train.py -- same CLI and epoch structure as the reference's train.py:49-158, with the hot path
(model, optimizer, style augmentor, epoch loop) served by speedplusbaseline_b200.  `--no_cuda` runs the
reference's unmodified torch modules (it needs --reference_root), exactly the role that flag has upstream."""
import json
import logging
import os

import torch

from config import cfg

logger = logging.getLogger(__name__)


def main():
    from speedplusbaseline_b200 import cli
    from speedplusbaseline_b200.utils import set_all_seeds, save_checkpoint
    from speedplusbaseline_b200 import dist as D
    device = cli.select_device(cfg)
    cli.setup_logger('train')
    logger.info('Random seed value: {}'.format(cfg.seed))
    set_all_seeds(cfg.seed + D.rank(), cfg, True)       # per-rank RNG streams (data order, style noise); weights are broadcast below
    os.makedirs(cfg.savedir, exist_ok=True)
    os.makedirs(cfg.logdir, exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        writer = SummaryWriter(cfg.logdir) if D.is_main() else None
    except Exception:                                   # tensorboard not installed: keep training
        writer = None
    with open(os.path.join(cfg.savedir, 'config.txt'), 'w') as f:
        json.dump(cfg.__dict__, f, indent=2)

    if device.type == 'cuda':
        from speedplusbaseline_b200.nets.build import get_model, get_optimizer
        from speedplusbaseline_b200.core import trainer as T
        from speedplusbaseline_b200.styleaug.styleAugmentor import StyleAugmentor
    else:                                               # --no_cuda: the reference's own path
        cli.reference_modules(cfg)
        from src.nets.build import get_model, get_optimizer
        import src.core.trainer as T
        from src.styleaug.styleAugmentor import StyleAugmentor

    model = get_model(cfg)
    D.broadcast_model(model)
    styleAugmentor = None
    if cfg.randomize_texture:
        styleAugmentor = StyleAugmentor(cfg.texture_alpha, device)
        logger.info('Texture randomization enabled with alpha = {}'.format(cfg.texture_alpha))
        logger.info('   - Randomization ratio: {:.2f}'.format(cfg.texture_ratio))
    optimizer = get_optimizer(cfg, model)
    begin_epoch = cli.resume(cfg, model, optimizer, device)
    best_perf = begin_epoch
    model = model.to(device)
    scaler = None
    if cfg.fp16:
        scaler = torch.amp.GradScaler('cuda', enabled=device.type == 'cuda')
        logger.info('Mixed-precision training enabled')
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=cfg.lr_decay_step, gamma=cfg.lr_decay_alpha)
    train_loader, test_loader = cli.make_loaders(cfg, [dict(is_train=True, is_source=True), dict(is_train=False, is_source=False)])
    train_epoch = getattr(T, 'train_single_epoch_' + cfg.model_name)

    for epoch in range(begin_epoch, cfg.max_epochs):
        train_epoch(epoch + 1, cfg, model, train_loader, optimizer, writer, device,
                    styleAugmentor=styleAugmentor, scaler=scaler)
        lr_scheduler.step()
        if cfg.test_epoch > 0 and (epoch + 1) % cfg.test_epoch == 0:
            cli.validate(cfg, model, test_loader, epoch + 1, writer, device)
        perf = epoch + 1
        is_best = perf > best_perf
        best_perf = max(best_perf, perf)
        if not D.is_main():                          # replicas are identical after every step: rank 0 writes the checkpoint
            continue
        save_checkpoint({'epoch': epoch + 1, 'model': cfg.model_name, 'state_dict': model.state_dict(),
                         'best_score': best_perf, 'optimizer': optimizer.state_dict()}, is_best, cfg.savedir)
    if writer is not None:
        writer.close()
    if D.world_size() > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
