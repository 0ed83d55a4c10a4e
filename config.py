"""Command-line configuration -- flag names, dests and defaults of the reference's config.py:13-61
(kept verbatim so train.py / adapt.py / test.py invocations carry over), plus three non-conflicting
additions for running without the SPEED+ dataset.  Parsed at import, like the reference (config.py:64)."""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))

_FLAGS = [
    # (flag, kwargs)                                                      reference line
    ('--seed', dict(type=int, default=2021)),                             # :13
    ('--projroot', dict(type=str, default=_HERE)),                        # :14 (author's path there)
    ('--dataroot', dict(type=str, default=os.path.join(_HERE, 'data'))),  # :15
    ('--dataname', dict(type=str, default='speedplus')),
    ('--savedir', dict(type=str, default='checkpoints/synthetic/krn')),
    ('--resultfn', dict(type=str, default='')),
    ('--logdir', dict(type=str, default='log/synthetic/krn')),
    ('--pretrained', dict(type=str, default='')),
    ('--model_name', dict(type=str, default='krn')),                      # :24
    ('--input_shape', dict(nargs='+', type=int, default=(224, 224))),
    ('--num_keypoints', dict(type=int, default=11)),
    ('--num_classes', dict(type=int, default=5000)),
    ('--num_neighbors', dict(type=int, default=5)),
    ('--keypts_3d_model', dict(type=str, default='src/utils/tangoPoints.mat')),
    ('--attitude_class', dict(type=str, default='src/utils/attitudeClasses.mat')),
    ('--start_over', dict(dest='auto_resume', action='store_false', default=True)),          # :34
    ('--randomize_texture', dict(dest='randomize_texture', action='store_true', default=False)),
    ('--perform_dann', dict(dest='dann', action='store_true', default=False)),
    ('--texture_alpha', dict(type=float, default=0.5)),
    ('--texture_ratio', dict(type=float, default=0.5)),
    ('--use_fp16', dict(dest='fp16', action='store_true', default=False)),
    ('--batch_size', dict(type=int, default=32)),
    ('--max_epochs', dict(type=int, default=75)),
    ('--num_workers', dict(type=int, default=8)),
    ('--test_epoch', dict(type=int, default=-1)),
    ('--optimizer', dict(type=str, default='rmsprop')),
    ('--lr', dict(type=float, default=0.001)),
    ('--momentum', dict(type=float, default=0.9)),
    ('--weight_decay', dict(type=float, default=5e-5)),
    ('--lr_decay_alpha', dict(type=float, default=0.96)),
    ('--lr_decay_step', dict(type=int, default=1)),
    ('--train_domain', dict(type=str, default='synthetic')),              # :52
    ('--test_domain', dict(type=str, default='lightbox')),
    ('--train_csv', dict(type=str, default='train.csv')),
    ('--test_csv', dict(type=str, default='lightbox.csv')),
    ('--gpu_id', dict(type=int, default=0)),                              # :60
    ('--no_cuda', dict(dest='use_cuda', action='store_false', default=True)),
    # ---- additions (do not exist upstream) ----
    ('--synthetic_data', dict(type=int, default=0, help='N>0: train on N synthetic iterations per epoch '
                                                        'instead of the SPEED+ loaders (datasets are out of scope)')),
    ('--reference_root', dict(type=str, default=os.environ.get('SPEEDPLUS_REFERENCE', ''),
                              help='checkout of tpark94/speedplusbaseline providing src.datasets / src.core.inference')),
    ('--no_graph', dict(dest='use_graph', action='store_false', default=True)),
    ('--device_transforms', dict(action='store_true', default=False,
                                 help='decode-only DataLoader workers + the transform stack (crop/resize/augment) on the GPU '
                                      '(speedplusbaseline_b200/datasets); KRN train/test and SPN test')),
]


def build_parser():
    p = argparse.ArgumentParser('Configurations for SPEED+ Baseline Study (B200-native hot path)')
    for flag, kw in _FLAGS:
        p.add_argument(flag, **kw)
    return p


parser = build_parser()
# pytest and other importers pass their own argv: only parse ours when run as a CLI
_cli = os.path.basename(sys.argv[0]) in ('train.py', 'adapt.py', 'test.py')
cfg = parser.parse_args(sys.argv[1:] if _cli else [])
