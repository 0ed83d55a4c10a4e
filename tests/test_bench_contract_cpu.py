"""Host-side contract checks that need no GPU: the reference arm of bench.py (the UNMODIFIED reference loop staged under
baseline/_ref), the staging manifest, the CLI device / process-group set-up for a single process."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'src', 'core', 'trainer.py'))


@pytest.mark.skipif(not STAGED, reason='baseline/_ref not staged (python tools/stage_reference.py)')
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'krn_train_images_per_sec' and line['unit'] == 'images/s'
    assert line['steps'] == 1 and line['warmup'] == 1 and line['higher_is_better'] is True
    cb = line['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and 'train_single_epoch_krn' in cb['sample'] and cb['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['value'] > 0 and abs(line['ms_per_step'] * line['value'] / 1e3 - 48) < 1e-6 * 48 + 1e-3


@pytest.mark.skipif(not STAGED, reason='baseline/_ref not staged')
def test_staged_reference_is_the_unmodified_checkout():
    import hashlib
    m = json.load(open(os.path.join(ROOT, 'baseline', '_ref', 'MANIFEST.json')))
    assert 'src/core/trainer.py' in m['files'] and 'src/styleaug/checkpoints/checkpoint_transformer.pth' in m['files']
    src = m['source']
    for rel, sha in m['files'].items():
        staged = os.path.join(ROOT, 'baseline', '_ref', rel)
        assert hashlib.sha256(open(staged, 'rb').read()).hexdigest() == sha, rel
        if os.path.isdir(src):                                   # build container: byte-identical to the reference checkout
            assert open(os.path.join(src, rel), 'rb').read() == open(staged, 'rb').read(), rel


def test_cli_device_setup_single_process(monkeypatch):
    from speedplusbaseline_b200 import dist as D
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    cfg = types.SimpleNamespace(use_cuda=False)
    dev = D.setup_cli(cfg)
    assert dev.type == 'cpu' and cfg.device is None and cfg.rank == 0 and cfg.world_size == 1
    assert D.world_size() == 1 and D.rank() == 0 and D.is_main()
    json.dumps(cfg.__dict__)                                     # train.py dumps cfg.__dict__ to config.txt


def test_torch_ops_schemas_registered_without_a_gpu():
    import torch
    import speedplusbaseline_b200.torch_ops as T
    for n in T.OPS:
        assert hasattr(torch.ops.b200sp, n), n
    with pytest.raises(NotImplementedError):                     # CUDA-only registration: no CPU fallback
        torch.ops.b200sp.conv1x1_dgrad(torch.randn(4, 8), torch.randn(8, 8))
