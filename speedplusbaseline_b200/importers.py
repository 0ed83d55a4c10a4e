"""Pretrained-weight importers: the two external weight files the reference consumes, converted
into reference-keyed state dicts that ParamStore.load_state_dict() lays out in the kernel-native flat buffer.

  * torchvision MobileNetV2 ImageNet weights -- park2019.py:107-108 builds `self.base` from
    `models.mobilenet_v2(pretrained=True).features[:-1]`, i.e. checkpoint key `features.<i>.<rest>` (i = 0..17) becomes
    `base.<i>.<rest>` (`net.base.<i>.<rest>` inside RevGrad, revgrad.py:64); `features.18.*` (the 1x1 320->1280 conv the
    reference drops) and `classifier.*` are not used.  torchvision 0.9 (requirements.txt:7) and current releases share
    these keys.
  * Caffe AlexNet dump `bvlc_alexnet.npy` -- spn.py:104-123: a pickled dict {layer: [W (H,W,Cin/groups,Cout), b (Cout,)]};
    only conv1..conv5 are read, W transposed to torch's [Cout, Cin/groups, H, W].

Pure host code (no CUDA needed): the engines' load_state_dict does the HBM layout conversion.
"""
import re
from collections import OrderedDict

import numpy as np
import torch

_FEAT = re.compile(r'^(?:module\.)?features\.(\d+)\.(.+)$')
N_BASE = 18                      # features[0..17] are kept (park2019.py:108 drops the last child)
ALEXNET_CONVS = ('conv1', 'conv2', 'conv3', 'conv4', 'conv5')


def mobilenetv2_to_krn(tv_state_dict, prefix=''):
    """torchvision `mobilenet_v2().state_dict()` -> {`<prefix>base.<i>....`: tensor} for the KRN / RevGrad backbone.
    Raises KeyError when the checkpoint does not hold all 18 feature blocks."""
    out = OrderedDict()
    seen = set()
    for k, v in tv_state_dict.items():
        m = _FEAT.match(k)
        if not m:
            continue                                  # classifier.*, or anything that is not a feature block
        i = int(m.group(1))
        if i >= N_BASE:
            continue                                  # features.18: dropped by the reference
        seen.add(i)
        out['%sbase.%d.%s' % (prefix, i, m.group(2))] = v
    if seen != set(range(N_BASE)):
        raise KeyError('not a torchvision MobileNetV2 state_dict: feature blocks %s missing' % sorted(set(range(N_BASE)) - seen))
    return out


def load_mobilenetv2_backbone(model, checkpoint):
    """Load ImageNet MobileNetV2 weights (path to `mobilenet_v2-*.pth` or a state dict) into a KeypointRegressionNet or
    RevGrad facade, leaving extras / head / domain classifier untouched.  Returns the list of keys written."""
    sd = torch.load(checkpoint, map_location='cpu', weights_only=True) if isinstance(checkpoint, (str, bytes)) else checkpoint
    store = model._store if hasattr(model, '_store') else model
    prefix = 'net.' if any(k.startswith('net.base.') for k in store.ref_keys()) else ''
    mapped = mobilenetv2_to_krn(sd, prefix)
    rk = store.ref_keys()
    bad = [k for k in mapped if k not in rk]
    if bad:
        raise KeyError('keys not present in the target model: %s' % bad[:6])
    store.load_state_dict(mapped, strict=False)
    return list(mapped)


def alexnet_npy_to_spn(weights_dict):
    """Caffe AlexNet dict -> {conv<i>.weight [Cout,Cin/g,H,W], conv<i>.bias} (spn.py:111-123); other layers are ignored."""
    sd = OrderedDict()
    for name, blobs in weights_dict.items():
        key = name.decode() if isinstance(name, bytes) else name
        if key not in ALEXNET_CONVS:
            continue
        for data in blobs:
            data = np.asarray(data)
            if data.ndim == 4:
                sd[key + '.weight'] = torch.from_numpy(np.ascontiguousarray(np.transpose(data, (3, 2, 0, 1)))).float()
            else:
                sd[key + '.bias'] = torch.from_numpy(np.ascontiguousarray(data)).float()
    return sd


def load_alexnet_npy(model, weight_path):
    """spn.py:104-123 for the SPN facade (or a bare ParamStore).  Returns the keys written."""
    weights_dict = np.load(weight_path, allow_pickle=True, encoding='bytes').item()
    sd = alexnet_npy_to_spn(weights_dict)
    store = model._store if hasattr(model, '_store') else model
    store.load_state_dict(sd, strict=False)
    return list(sd)
