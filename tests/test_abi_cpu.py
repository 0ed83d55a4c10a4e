"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200sp.h declares;
host-side layout logic (ParamStore) round-trips the reference's state_dict.  No compute calls."""
import os
import re

import torch

from oracle import krn as okrn, revgrad as orev, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'b200sp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200sp_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from speedplusbaseline_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), 'libb200sp.so does not export %s' % n
    assert _lib.lib.b200sp_version() >= 1


def test_binding_covers_header():
    from speedplusbaseline_b200 import _lib
    assert set(_declared()) <= set(_lib.EXPORTS) | {'b200sp_colsum_f32'}


def test_param_store_roundtrip_krn_cpu():
    from speedplusbaseline_b200.krn_engine import krn_layout
    from speedplusbaseline_b200.params import ParamStore
    W, BN, order = krn_layout(11)
    st = ParamStore(W, BN, torch.device('cpu'))
    sd = synth.synth_state_dict(okrn.krn_shapes(), 3)
    assert order == list(sd.keys())                      # same keys, same order as the reference module
    st.load_state_dict(sd)
    back = st.state_dict(order)
    for k in sd:
        assert back[k].shape == sd[k].shape and torch.equal(back[k], sd[k]), k
    # native layouts: depthwise [9][C], head [N][7][7][C]
    dw = st.view('base.1.conv.0.0.weight').view(9, 32)
    assert torch.equal(dw[4], sd['base.1.conv.0.0.weight'][:, 0, 1, 1])
    hw = st.view('head.0.weight').view(22, 7, 7, 1024)
    assert torch.equal(hw[3, 2, 5], sd['head.0.weight'][3, :, 2, 5])


def test_param_store_roundtrip_revgrad_cpu():
    from speedplusbaseline_b200.krn_engine import krn_layout
    from speedplusbaseline_b200.params import ParamStore
    W, BN, order = krn_layout(11, prefix='net.', dann=True)
    st = ParamStore(W, BN, torch.device('cpu'))
    sd = synth.synth_state_dict(orev.revgrad_shapes(), 5)
    assert order == list(sd.keys())
    st.load_state_dict(sd)
    back = st.state_dict(order)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_strict_load_reports_missing_keys():
    import pytest
    from speedplusbaseline_b200.krn_engine import krn_layout
    from speedplusbaseline_b200.params import ParamStore
    W, BN, order = krn_layout(11)
    st = ParamStore(W, BN, torch.device('cpu'))
    sd = synth.synth_state_dict(okrn.krn_shapes(), 3)
    sd.pop('head.0.bias')
    with pytest.raises(RuntimeError):
        st.load_state_dict(sd, strict=True)
