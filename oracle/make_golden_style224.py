"""tests/golden/styleaug_real_224.npz: the UNMODIFIED reference StyleAugmentor (styleAugmentor.py:12-68) with its REAL
checkpoints (Ghiasi transformer weights, PBN embedding statistics, SPEED+ mean embedding) on one seeded 224x224 image --
BASELINE.json's resolution.  Build container only (needs /root/reference); same shims as oracle/make_golden.py.

    python -m oracle.make_golden_style224
"""
import os

import numpy as np


def main():
    from oracle.make_golden import _shims, OUT
    _shims()
    import torch
    from src.styleaug.styleAugmentor import StyleAugmentor
    from oracle import synth
    torch.set_num_threads(8)
    aug = StyleAugmentor(0.5, torch.device('cpu'))
    x = synth.synth_images(1, 224, 224, seed=7)
    torch.manual_seed(123)
    noise = torch.randn(1, 100)
    torch.manual_seed(123)
    out = aug(x)
    # values lie in (0,1): float16 keeps 5e-4 absolute, an order below the parity gate, and halves the fixture
    np.savez_compressed(os.path.join(OUT, 'styleaug_real_224.npz'), out=out.numpy().astype(np.float16), noise=noise.numpy(),
                        x_sum=synth.checksum(x))
    print('styleaug_real_224.npz', out.shape, float(out.mean()))


if __name__ == '__main__':
    main()
