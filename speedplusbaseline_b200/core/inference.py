"""Batched evaluation with the post-processing tail on the device.

The reference evaluates one image at a time and moves every logit to the host (`xc.cpu(), yc.cpu()`,
park2019.py:165; `topClasses[b].cpu()`, /root/reference/src/core/inference.py:185).  Here a whole batch goes through
ONE forward of the B200 engine, the tensor part of the tail runs as a libb200sp launch

    KRN  b200sp_kpt_denorm    keypoints in the crop frame -> pixels            (inference.py:236-243)
    SPN  b200sp_topk_softmax  top-`num_neighbors` attitude classes + softmax   (inference.py:180-181)

and one small pinned device->host copy per batch hands the result to the CPU pose code.  EPnP, the quaternion mean, the
SPN position solve and the SPEED metrics are CPU post-processing outside the hot path: `valid_krn` /
`valid_spn` (same signatures as inference.py:43,148) take them from the reference checkout (`cfg.reference_root`).
There is no CPU fallback for the device part.
"""
import os.path as osp
import time

import numpy as np
import torch

from .. import _lib as L
from ..utils import AverageMeter, report_progress


def _device_images(model, images):
    dev = model.engine.device
    return images.to(dev, non_blocking=True).contiguous().float()


@torch.no_grad()
def krn_keypoints_pix(model, images, bbox):
    """images [B,3,224,224] in [0,1], bbox [B,4] = (xmin, xmax, ymin, ymax) pixels -> host tensor [B,K,2] of keypoint
    pixel coordinates (what inference.py:_keypts_to_pose feeds to pnp), eval-mode forward."""
    eng = model.engine
    x = _device_images(model, images)
    cx = eng.forward(x, None, train=False)
    B, K = x.shape[0], eng.N // 2
    bb = bbox.to(eng.device, non_blocking=True).contiguous().float()
    out = torch.empty(B, K, 2, device=eng.device)
    L.call('b200sp_kpt_denorm', cx.logits.data_ptr(), bb.data_ptr(), out.data_ptr(), B, K, L.stream_ptr())
    host = torch.empty(B, K, 2, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host


@torch.no_grad()
def spn_top_classes(model, images, k):
    """images [B,3,227,227] -> (topWeights [B,k] softmaxed, topClasses [B,k] int64) on the host: the regress-branch
    logits ("weights", spn.py:143) never leave the device."""
    eng = model.engine
    x = _device_images(model, images)
    _, r = eng.forward(x, train=False)
    B, N = r.shape
    tw = torch.empty(B, k, device=eng.device)
    ti = torch.empty(B, k, dtype=torch.int64, device=eng.device)
    L.call('b200sp_topk_softmax', r.data_ptr(), tw.data_ptr(), None, ti.data_ptr(), B, N, k, L.stream_ptr())
    hw, hi = torch.empty(B, k, pin_memory=True), torch.empty(B, k, dtype=torch.int64, pin_memory=True)
    hw.copy_(tw, non_blocking=True)
    hi.copy_(ti, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return hw, hi


class _Meters:
    NAMES = (('eR', 'deg'), ('eT', 'm'), ('speed (raw)', '-'), ('speed (thr)', '-'))

    def __init__(self):
        self.time, self.acc = AverageMeter('ms'), AverageMeter('%')
        self.m = {n: AverageMeter(u) for n, u in self.NAMES}
        self.rows = {n: [] for n, _ in self.NAMES}

    def add(self, vals, acc):
        for (n, _), v in zip(self.NAMES, vals):
            self.m[n].update(v, 1)
            self.rows[n].append(v)
        self.acc.update(acc * 100, 1)


def _validate(epoch, cfg, model, data_loader, writer, poses_of_batch, ref):
    """shared driver: `poses_of_batch(images, bbox)` -> list of (q_pr, t_pr); `ref` = the reference's CPU metric functions."""
    mt = _Meters()
    model.eval()
    for idx, (images, bbox, q_gt, t_gt) in enumerate(data_loader):
        start = time.time()
        poses = poses_of_batch(images, bbox)
        for b, (q_pr, t_pr) in enumerate(poses):
            q_i, t_i = q_gt[b].numpy(), t_gt[b].numpy()
            raw, acc = ref['speed_score'](t_pr, q_pr, t_i, q_i, applyThresh=False)
            thr, _ = ref['speed_score'](t_pr, q_pr, t_i, q_i, applyThresh=True, rotThresh=0.169, posThresh=0.002173)
            mt.add((ref['error_orientation'](q_pr, q_i), ref['error_translation'](t_pr, t_i), raw, thr), acc)
        mt.time.update((time.time() - start) * 1000, len(poses))
        report_progress(epoch=epoch, lr=np.nan, epoch_iter=idx + 1, epoch_size=len(data_loader), time=mt.time, is_train=False,
                        eT=mt.m['eT'], eR=mt.m['eR'], speed=mt.m['speed (raw)'], acc=mt.acc)
    if writer is not None:
        for tag, n in (('Valid/err_q [deg]', 'eR'), ('Valid/err_t [m]', 'eT'), ('Valid/speed (raw) [-]', 'speed (raw)'),
                       ('Valid/speed (thr) [-]', 'speed (thr)')):
            writer.add_scalar(tag, mt.m[n].avg, epoch)
    return mt


def _reference_cpu_tail(cfg):
    """EPnP / quaternion mean / SPN position / SPEED metrics: CPU code of the reference checkout (out of scope)."""
    from ..cli import reference_modules
    reference_modules(cfg)
    from src.utils.utils import pnp, weighted_mean_quaternion
    from src.utils.computePositionSPN import compute_position_spn
    from src.utils import metrics
    return dict(pnp=pnp, weighted_mean_quaternion=weighted_mean_quaternion, compute_position_spn=compute_position_spn,
                error_orientation=metrics.error_orientation, error_translation=metrics.error_translation,
                speed_score=metrics.speed_score)


def valid_krn(epoch, cfg, model, data_loader, cameraMatrix, distCoeffs, corners3D, writer, device, qClass=None, ref=None):
    """inference.py:43-146 with a batched forward and on-device keypoint de-normalisation."""
    ref = ref or _reference_cpu_tail(cfg)

    def poses(images, bbox):
        pix = krn_keypoints_pix(model, images, bbox).numpy()
        return [ref['pnp'](corners3D, pix[b], cameraMatrix, distCoeffs) for b in range(pix.shape[0])]

    mt = _validate(epoch, cfg, model, data_loader, writer, poses, ref)
    for fn, n in (('err_q.txt', 'eR'), ('err_t.txt', 'eT'), ('speed_raw.txt', 'speed (raw)'), ('speed_mod.txt', 'speed (thr)')):
        with open(osp.join(cfg.logdir, fn), 'w') as f:
            f.writelines('{:.5f}\n'.format(v) for v in mt.rows[n])
    return mt.m


def valid_spn(epoch, cfg, model, data_loader, cameraMatrix, distCoeffs, corners3D, writer, device, qClass, ref=None):
    """inference.py:148-221 with a batched forward and on-device top-k + softmax."""
    ref = ref or _reference_cpu_tail(cfg)

    def poses(images, bbox):
        tw, ti = spn_top_classes(model, images, cfg.num_neighbors)
        out = []
        for b in range(tw.shape[0]):
            q_pr = ref['weighted_mean_quaternion'](qClass[ti[b].numpy()], tw[b].numpy())
            out.append((q_pr, ref['compute_position_spn'](q_pr, bbox[b].numpy(), corners3D, cameraMatrix, distCoeffs)))
        return out

    return _validate(epoch, cfg, model, data_loader, writer, poses, ref).m
