#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout under baseline/_ref/ so that it travels to the GPU box.

baseline/_ref/ is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored,
so `bench.py --impl reference`, the `reference_cuda_*` bench secondaries and the real-checkpoint style-aug
tests can import it on a box where /root/reference does not exist.  `pip install /root/reference` is not
possible (no setup.py / pyproject.toml), hence a plain file copy: every *.py file plus the small data files
the hot path loads (style-transfer weights, embedding statistics, attitude classes) -- ~8.2 MB.  The 84 MB of
StylePredictor / raw-embedding files are not on the path (SURVEY.md section 2 "out of scope") and are skipped.

    python tools/stage_reference.py            # /root/reference -> baseline/_ref
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('B200SP_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
DATA = ('src/styleaug/checkpoints/checkpoint_transformer.pth', 'src/styleaug/checkpoints/checkpoint_embeddings.pth',
        'src/styleaug/checkpoints/embedding_mean_speedplus.npy', 'src/utils/attitudeClasses.mat', 'src/utils/tangoPoints.mat')


def _sha(p):
    h = hashlib.sha256()
    with open(p, 'rb') as f:
        for blk in iter(lambda: f.read(1 << 20), b''):
            h.update(blk)
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print('reference checkout %s not present: nothing staged' % SRC)
        return None
    manifest = {}
    for dp, dn, fn in os.walk(SRC):
        dn[:] = [d for d in dn if d not in ('.git', '__pycache__')]
        for f in fn:
            rel = os.path.relpath(os.path.join(dp, f), SRC)
            if not (f.endswith('.py') or rel in DATA or f in ('requirements.txt', 'LICENSE.md', 'LICENSE')):
                continue
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(dp, f), out)
            manifest[rel] = _sha(out)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SRC, 'files': manifest}, f, indent=1, sort_keys=True)
    if verbose:
        tot = sum(os.path.getsize(os.path.join(DST, r)) for r in manifest)
        print('staged %d files (%.1f MB) from %s to %s' % (len(manifest), tot / 1e6, SRC, DST))
    return DST


if __name__ == '__main__':
    sys.exit(0 if stage() or not os.path.isdir(SRC) else 1)
