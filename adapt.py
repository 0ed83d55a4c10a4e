#!/usr/bin/env python
"""adapt.py -- DANN domain adaptation with the reference's CLI (adapt.py:47-146): requires
--perform_dann and --model_name krn, fixed seed 2021, fp32 only (adapt.py:99-101)."""
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
    assert cfg.dann and cfg.model_name == 'krn'
    set_all_seeds(2021 + D.rank(), cfg, True)           # adapt.py:53 fixes 2021; ranks > 0 offset it (weights are broadcast below)
    cli.setup_logger('train')
    os.makedirs(cfg.savedir, exist_ok=True)
    os.makedirs(cfg.logdir, exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        writer = SummaryWriter(cfg.logdir) if D.is_main() else None
    except Exception:
        writer = None
    with open(os.path.join(cfg.savedir, 'config.txt'), 'w') as f:
        json.dump(cfg.__dict__, f, indent=2)
    if device.type == 'cuda':
        from speedplusbaseline_b200.nets.build import get_model, get_optimizer
        from speedplusbaseline_b200.core.dann import train_dann_single_epoch_krn
    else:
        cli.reference_modules(cfg)
        from src.nets.build import get_model, get_optimizer
        from src.core.dann import train_dann_single_epoch_krn
    model = get_model(cfg)
    D.broadcast_model(model)
    optimizer = get_optimizer(cfg, model)
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=cfg.lr_decay_step, gamma=cfg.lr_decay_alpha)
    begin_epoch = cli.resume(cfg, model, optimizer, device)
    best_perf = begin_epoch
    model.to(device)
    src_loader, tgt_loader, test_loader = cli.make_loaders(cfg, [
        dict(is_train=True, is_source=True, load_labels=True),
        dict(is_train=True, is_source=False, load_labels=False),
        dict(is_train=False, is_source=False, load_labels=True)])
    for epoch in range(begin_epoch, cfg.max_epochs):
        train_dann_single_epoch_krn(epoch, cfg, model, src_loader, tgt_loader, optimizer, writer, device)
        lr_scheduler.step()
        if cfg.test_epoch > 0 and (epoch + 1) % cfg.test_epoch == 0:
            cli.validate(cfg, model, test_loader, epoch, writer, device)
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
