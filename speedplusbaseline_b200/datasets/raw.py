"""Raw-frame datasets + a loader wrapper that runs the transform stack on the device.

The reference's datasets (Park2019KRNDataset.py:44-109, SPNDataset.py) decode one image, run the PIL / torch transform stack on
it inside a DataLoader worker and collate float tensors.  Here the workers only DECODE (`RawFrameDataset`: the same CSV
columns and split logic, the 8-bit frame handed over untouched) and `DeviceBatchLoader` applies `DeviceTransforms` once per
collated batch on the GPU, yielding exactly what the reference loaders yield:

    train, labels   (images [B,3,h,w], keypts [B,2,K])          Park2019KRNDataset.py:103-105
    train, no labels images                                     :106-107  (DANN target domain)
    test            (images, bbox [B,4], q_gt [B,4], t_gt [B,3]) :108-111

so `train_single_epoch_krn` / `train_dann_single_epoch_krn` / `valid_krn` consume it unchanged (their DevicePrefetcher passes
device tensors through).  SPN soft attitude-class targets (SPNDataset.py:83-94) depend on the attitude-class table and stay with
the reference's dataset: for `model_name == 'spn'` only the test-time path (ResizeCrop) is provided here.
"""
import os.path as osp

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .transforms import build_transforms


class RawFrameDataset(Dataset):
    """CSV rows: image path | xmin xmax ymin ymax | q0..q3 t1..t3 | kx1 ky1 .. kx11 ky11  (Park2019KRNDataset.py:38-43)."""

    def __init__(self, cfg, is_train=True, is_source=True, load_labels=True):
        import pandas as pd
        self.is_train, self.load_labels = is_train, load_labels
        self.root = osp.join(cfg.dataroot, cfg.dataname)
        self.num_keypts = cfg.num_keypoints
        if is_train and is_source:
            assert load_labels
            csvfile = osp.join(self.root, cfg.train_domain, 'splits_' + cfg.model_name, cfg.train_csv)
        else:
            assert is_train is False or not load_labels        # target-domain training images carry no labels (:62-63)
            csvfile = osp.join(self.root, cfg.test_domain, 'splits_' + cfg.model_name, cfg.test_csv)
        self.csv = pd.read_csv(csvfile, header=None)

    def __len__(self):
        return len(self.csv)

    def __getitem__(self, index):
        from PIL import Image
        assert index < len(self), 'Index range error'
        row = self.csv.iloc[index]
        img = Image.open(osp.join(self.root, row[0]))
        if img.mode != 'L':                      # SPEED+ frames are 8-bit grey; anything else goes through RGB like upstream
            img = img.convert('RGB')
        frame = torch.from_numpy(np.array(img, dtype=np.uint8))
        bbox = torch.from_numpy(np.array(row[1:5], dtype=np.float32))
        if self.is_train and self.load_labels:
            k = np.array(row[12:], dtype=np.float32)
            keypts = torch.from_numpy(np.ascontiguousarray(np.transpose(np.reshape(k, (self.num_keypts, 2)))))     # [2, K] pixels
        else:
            keypts = torch.zeros(2, self.num_keypts)
        q_gt = torch.from_numpy(np.array(row[5:9], dtype=np.float32))
        t_gt = torch.from_numpy(np.array(row[9:12], dtype=np.float32))
        return frame, bbox, keypts, q_gt, t_gt


class DeviceBatchLoader:
    """wraps a loader of collated raw samples (frames uint8 [B,H,W(,C)], bbox, keypts_pix, q_gt, t_gt)."""

    def __init__(self, raw_loader, transforms, is_train=True, load_labels=True):
        self.raw, self.tf, self.is_train, self.load_labels = raw_loader, transforms, is_train, load_labels

    def __len__(self):
        return len(self.raw)

    def __iter__(self):
        for frames, bbox, keypts, q_gt, t_gt in self.raw:
            images, bbox_out, k = self.tf(frames, bbox.numpy(), keypts.numpy())
            if self.is_train:
                yield (images, k) if self.load_labels else images
            else:
                yield images, bbox_out, q_gt, t_gt


def make_dataloader(cfg, is_train=True, is_source=True, load_labels=True, device=None, generator=None):
    """src/datasets/build.py:45-66 with the transform stack moved behind the collate."""
    ds = RawFrameDataset(cfg, is_train, is_source, load_labels)
    raw = DataLoader(ds, batch_size=cfg.batch_size if is_train else 1, shuffle=is_train,
                     num_workers=cfg.num_workers if is_train else 1, pin_memory=True, drop_last=True)
    tf = build_transforms(cfg.model_name, cfg.input_shape, p_aug=0.5, is_train=is_train, device=device, generator=generator)
    return DeviceBatchLoader(raw, tf, is_train, load_labels)
