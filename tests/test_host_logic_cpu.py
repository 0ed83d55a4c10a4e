"""Host-side logic that needs no GPU: CLI flag parity with the reference's config.py, the parameter store's
layout conversions, the shifted-GEMM chunk planner of the style net (emulated on the CPU against conv2d), schedules."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_config_flags_match_reference():
    """every flag of the reference's config.py:13-61 exists here with the same dest and default (paths aside)"""
    if not os.path.exists(os.path.join(REF, 'config.py')):
        pytest.skip('reference checkout only exists in the build container')
    import importlib.util
    argv = sys.argv
    sys.argv = ['x']
    try:
        spec = importlib.util.spec_from_file_location('ref_config', os.path.join(REF, 'config.py'))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        sys.path.insert(0, ROOT)
        import config as ours
    finally:
        sys.argv = argv
    ref_actions = {a.option_strings[0]: a for a in ref.parser._actions if a.option_strings and a.option_strings[0] != '-h'}
    our_actions = {a.option_strings[0]: a for a in ours.parser._actions if a.option_strings and a.option_strings[0] != '-h'}
    assert set(ref_actions) <= set(our_actions), set(ref_actions) - set(our_actions)
    for flag, a in ref_actions.items():
        b = our_actions[flag]
        assert a.dest == b.dest, flag
        if flag not in ('--projroot', '--dataroot'):
            assert a.default == b.default, (flag, a.default, b.default)
        assert type(a) is type(b), flag


def test_param_store_layouts_roundtrip_on_cpu():
    from speedplusbaseline_b200.params import ParamStore
    from speedplusbaseline_b200.krn_engine import krn_layout
    from speedplusbaseline_b200.spn_engine import spn_layout
    from oracle import krn as okrn, spn as ospn, synth
    W, BN, order = krn_layout(11)
    st = ParamStore(W, BN, torch.device('cpu'))
    sd = synth.synth_state_dict(okrn.krn_shapes(), 3)
    st.load_state_dict(sd, strict=True)
    back = st.state_dict(order)
    assert list(back.keys()) == list(sd.keys())
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    assert all(e.off % 8 == 0 for e in st.entries.values())          # 16-byte aligned in fp32 and in the bf16 mirror
    W, order = spn_layout(40)
    st = ParamStore(W, [], torch.device('cpu'))
    shapes = ospn.spn_shapes(40)
    sd = {k: torch.randn(s) for k, s in shapes.items()}
    st.load_state_dict(sd, strict=True)
    back = st.state_dict(order)
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    # fc6 columns are stored in NHWC order: column (h*6+w)*256 + c of the native matrix == reference column c*36 + h*6 + w
    e = st.entries['fc6.weight']
    nat = st.params[e.off:e.off + e.numel].view(4096, 6, 6, 256)
    assert torch.equal(nat[5, 2, 3, 7], sd['fc6.weight'][5, 7 * 36 + 2 * 6 + 3])
    with pytest.raises(RuntimeError):
        st.load_state_dict({'nope': torch.zeros(1)}, strict=True)


@pytest.mark.parametrize('k,stride,Ci,Cp,Co', [(9, 1, 3, 8, 4), (3, 2, 32, 32, 8), (3, 2, 64, 64, 8), (3, 1, 128, 128, 16), (9, 1, 32, 32, 3)])
def test_shifted_gemm_plan_equals_conv2d(k, stride, Ci, Cp, Co):
    """CPU emulation of convtc.cu's data path from the host-side plan: planes (reflection-padded, phase-split),
    chunk list (plane, c0, row shift), packed weights -> the same numbers as conv2d on the padded input."""
    from speedplusbaseline_b200.styleaug.ghiasi import plan_chunks, pack_weight
    torch.manual_seed(0)
    B, H, W = 2, 12, 16
    pad = k // 2
    x = torch.randn(B, Ci, H, W, dtype=torch.float64)
    w = torch.randn(Co, Ci, k, k, dtype=torch.float64)
    xp = F.pad(x, (pad,) * 4, mode='reflect')
    ref = F.conv2d(xp, w, stride=stride)
    Ho, Wo = ref.shape[2], ref.shape[3]
    ps = stride
    Hq, Wq = xp.shape[2] // ps, xp.shape[3] // ps
    planes = []
    for qy in range(ps):
        for qx in range(ps):
            pl = torch.zeros(B * Hq * Wq + k * Wq + 64, Cp, dtype=torch.float64)   # rows past the tensor read as zero (TMA OOB fill)
            pl[:B * Hq * Wq, :Ci] = xp[:, :, qy::ps, qx::ps].permute(0, 2, 3, 1).reshape(-1, Ci)
            planes.append(pl)
    chunks, cols = plan_chunks(k, ps, Ci, Cp, Wq)
    seen = [e for ent in cols for e in ent if e is not None]
    assert len(seen) == len(set(seen)) == k * k * Ci and all(len(ent) == 64 for ent in cols)
    wm = pack_weight(w.float(), cols, 16).double()[:Co]                 # bf16-rounded weights
    wq = w.float().bfloat16().double()
    ref = F.conv2d(xp, wq, stride=stride)
    cbox = min(Cp, 64)
    P = 64 // cbox
    R = B * Hq * Wq
    acc = torch.zeros(R, Co, dtype=torch.float64)
    for j, (pl, c0, sh) in enumerate(chunks):
        flat = planes[pl].reshape(-1)
        # a 128-byte operand row = P consecutive pixels x cbox channels starting at pixel m + shift, channel c0
        rows = torch.stack([flat[(m + sh) * Cp + c0:(m + sh) * Cp + c0 + 64] if P == 1 else flat[(m + sh) * Cp:(m + sh) * Cp + 64] for m in range(R)])
        acc += rows @ wm[:, j * 64:(j + 1) * 64].t()
    got = acc.view(B, Hq, Wq, Co)[:, :Ho, :Wo].permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_dann_alpha_schedule_matches_oracle():
    from speedplusbaseline_b200.core.dann import dann_alpha
    from oracle.revgrad import dann_alpha as ref
    for idx, ep in ((0, 0), (3, 1), (99, 74)):
        assert dann_alpha(idx, ep, 100, 75) == ref(idx, ep, 100, 75)


def test_get_optimizer_dispatch_and_unknown_model_asserts():
    import types
    from speedplusbaseline_b200.nets import build
    with pytest.raises(AssertionError):
        build.get_model(types.SimpleNamespace(model_name='resnet', dann=False))


def test_ghiasi_param_table_matches_oracle_and_synthetic_state():
    """styleaug/ghiasi.py:param_shapes (product) lists exactly the reference module's 84 tensors (oracle table, which is
    pinned to the real checkpoint keys by the golden tests); synthetic_state feeds StyleAugmentor(state=...)."""
    from oracle import ghiasi as og
    from speedplusbaseline_b200.styleaug import ghiasi as pg
    a, b = pg.param_shapes(), og.ghiasi_shapes()
    assert set(a) == set(b) and all(tuple(a[k]) == tuple(b[k]) for k in b) and len(a) == 84
    st = pg.synthetic_state(3)
    assert set(st) == {'ghiasi', 'mean', 'cov', 'base'} and st['cov'].shape == (100, 100)
    assert all(tuple(st['ghiasi'][k].shape) == tuple(a[k]) for k in a)


def test_make_loaders_routes_spn_training_to_the_reference_dataset(monkeypatch):
    """--device_transforms: decode-only loaders for everything except SPN training (soft targets come from SPNDataset)."""
    import types
    from speedplusbaseline_b200 import cli
    calls = []
    fake_raw = types.ModuleType('speedplusbaseline_b200.datasets.raw')
    fake_raw.make_dataloader = lambda cfg, device=None, **s: calls.append(('device', s)) or 'dev'
    monkeypatch.setitem(sys.modules, 'speedplusbaseline_b200.datasets.raw', fake_raw)
    fake_ref = types.ModuleType('src.datasets.build')
    fake_ref.make_dataloader = lambda cfg, **s: calls.append(('reference', s)) or 'ref'
    for name in ('src', 'src.datasets'):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, 'src.datasets.build', fake_ref)
    monkeypatch.setattr(cli, 'reference_modules', lambda cfg: None)
    monkeypatch.setattr(cli.torch.cuda, 'is_available', lambda: True)
    train, test = dict(is_train=True, is_source=True, load_labels=True), dict(is_train=False, is_source=False, load_labels=True)
    cfg = types.SimpleNamespace(synthetic_data=0, device_transforms=True, use_cuda=True, model_name='spn')
    assert cli.make_loaders(cfg, [train, test]) == ['ref', 'dev']
    cfg.model_name = 'krn'
    assert cli.make_loaders(cfg, [train, test]) == ['dev', 'dev']
    cfg.device_transforms = False
    assert cli.make_loaders(cfg, [train, test]) == ['ref', 'ref']
