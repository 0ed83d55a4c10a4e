"""The reference CLIs (train.py / adapt.py / test.py, config.py flags) run end to end on the B200 path with
synthetic loaders: checkpoint files appear with the reference's names/keys and auto-resume continues."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, args, cwd):
    env = dict(os.environ, B200SP_QUIET='1', PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return r


def test_train_cli_krn_checkpoint_and_resume(tmp_path):
    common = ['--model_name', 'krn', '--optimizer', 'adamw', '--batch_size', '4', '--synthetic_data', '3', '--savedir', 'ck', '--logdir', 'lg',
              '--weight_decay', '0.01', '--lr_decay_alpha', '0.95']
    _run('train.py', common + ['--max_epochs', '1', '--start_over'], str(tmp_path))
    ck = torch.load(str(tmp_path / 'ck' / 'checkpoint.pth.tar'), map_location='cpu', weights_only=False)
    assert ck['epoch'] == 1 and ck['model'] == 'krn' and 'optimizer' in ck
    assert 'base.0.0.weight' in ck['state_dict'] and ck['state_dict']['head.0.weight'].shape == (22, 1024, 7, 7)
    assert int(ck['state_dict']['base.0.1.num_batches_tracked']) == 3
    _run('train.py', common + ['--max_epochs', '2'], str(tmp_path))          # auto-resume: one more epoch
    ck2 = torch.load(str(tmp_path / 'ck' / 'checkpoint.pth.tar'), map_location='cpu', weights_only=False)
    assert ck2['epoch'] == 2 and int(ck2['state_dict']['base.0.1.num_batches_tracked']) == 6
    best = torch.load(str(tmp_path / 'ck' / 'model_best.pth.tar'), map_location='cpu', weights_only=False)
    assert list(best.keys()) == list(ck2['state_dict'].keys())
    r = _run('test.py', ['--model_name', 'krn', '--batch_size', '4', '--synthetic_data', '2', '--pretrained', 'ck/model_best.pth.tar'], str(tmp_path))
    assert 'images/s' in r.stdout


def test_adapt_cli_dann(tmp_path):
    _run('adapt.py', ['--perform_dann', '--model_name', 'krn', '--optimizer', 'adamw', '--batch_size', '4', '--synthetic_data', '2',
                      '--max_epochs', '1', '--savedir', 'ck', '--logdir', 'lg', '--start_over'], str(tmp_path))
    ck = torch.load(str(tmp_path / 'ck' / 'checkpoint.pth.tar'), map_location='cpu', weights_only=False)
    assert 'net.base.0.0.weight' in ck['state_dict'] and 'domain_classifier.3.bias' in ck['state_dict']
    assert int(ck['state_dict']['net.base.0.1.num_batches_tracked']) == 4       # two forwards per iteration


def test_train_cli_with_style_augmentation_flag(tmp_path):
    """--randomize_texture: needs the style checkpoints; with none staged the CLI must fail loudly, not silently skip."""
    env = dict(os.environ, B200SP_QUIET='1', PYTHONPATH=ROOT)
    env.pop('SPEEDPLUS_STYLE_CKPT', None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'train.py'), '--model_name', 'krn', '--optimizer', 'adamw', '--batch_size', '4',
                        '--synthetic_data', '1', '--max_epochs', '1', '--savedir', 'ck', '--logdir', 'lg', '--start_over', '--randomize_texture'],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    from speedplusbaseline_b200.styleaug import styleAugmentor as SA
    try:
        SA.checkpoint_dir()
        have = True
    except FileNotFoundError:
        have = False
    assert (r.returncode == 0) == have, r.stderr[-2000:]
    if not have:
        assert 'checkpoints not found' in r.stderr


def test_train_cli_spn(tmp_path):
    _run('train.py', ['--model_name', 'spn', '--optimizer', 'adamw', '--batch_size', '4', '--synthetic_data', '2', '--max_epochs', '1',
                      '--input_shape', '227', '227', '--savedir', 'ck', '--logdir', 'lg', '--start_over'], str(tmp_path))
    ck = torch.load(str(tmp_path / 'ck' / 'checkpoint.pth.tar'), map_location='cpu', weights_only=False)
    assert ck['model'] == 'spn' and ck['state_dict']['fc6.weight'].shape == (4096, 9216)
    assert ck['state_dict']['conv1.weight'].shape == (96, 3, 11, 11)
    assert all(torch.isfinite(v).all() for v in ck['state_dict'].values())
