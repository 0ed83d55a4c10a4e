"""Host-side helpers with the reference's names (src/utils/utils.py:44-135, 289-299)."""
import logging
import os
import random
import sys

import numpy as np
import torch

logger = logging.getLogger(__name__)


class AverageMeter(object):
    """running average (utils.py:44-61)."""

    def __init__(self, unit='-'):
        self.unit = unit
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count else 0


def report_progress(epoch, lr, epoch_iter, epoch_size, time, is_train=True, **kwargs):
    """one-line progress bar (utils.py:81-105)."""
    if os.environ.get('B200SP_QUIET'):
        return
    frac = epoch_iter / max(1, epoch_size)
    bar = '#' * int(30 * frac) + ' ' * (30 - int(30 * frac))
    msg = '\r%s %03d (lr: %.5f): %04d/%04d [%s %3d%%] [%d (%d) ms] ' % (
        'Training' if is_train else 'Testing ', epoch, lr, epoch_iter, epoch_size, bar, int(100 * frac),
        time.val, time.avg)
    for k, m in kwargs.items():
        msg += '%s: %.2f (%.2f) [%s] ' % (k, m.val, m.avg, m.unit)
    sys.stdout.write(msg)
    sys.stdout.flush()
    if epoch_iter == epoch_size:
        sys.stdout.write('\n')
        sys.stdout.flush()


def save_checkpoint(states, is_best, output_dir, filename='checkpoint.pth.tar'):
    torch.save(states, os.path.join(output_dir, filename))
    if is_best and 'state_dict' in states:
        torch.save(states['state_dict'], os.path.join(output_dir, 'model_best.pth.tar'))


def load_checkpoint(checkpoint_file, model, optimizer, device):
    load_dict = torch.load(checkpoint_file, map_location='cpu', weights_only=False)
    model.load_state_dict(load_dict['state_dict'], strict=True)
    if optimizer is not None:
        optimizer.load_state_dict(load_dict['optimizer'])
    return load_dict['epoch'], load_dict['best_score']


def set_all_seeds(seed, cfg=None, use_cuda=True):
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if use_cuda and torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def num_total_parameters(model):
    return sum(p.numel() for p in model.parameters())


def num_trainable_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
