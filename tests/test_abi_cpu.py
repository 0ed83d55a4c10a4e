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


def _prototypes():
    """name -> list of parameter declarations, parsed from include/b200sp.h (comments stripped)."""
    src = open(os.path.join(ROOT, 'include', 'b200sp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|long long)\s+(b200sp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        params = [p.strip() for p in m.group(2).replace('\n', ' ').split(',')]
        out[m.group(1)] = [] if params in ([''], ['void']) else params
    return out


def test_ctypes_signatures_match_header_prototypes():
    """ABI drift guard: a ctypes argtypes list that disagrees with the C prototype corrupts the call silently on the GPU box.
    Every binding must have the prototype's parameter count, and pointer / integer / float kinds must line up."""
    import ctypes as C
    from speedplusbaseline_b200 import _lib
    protos = _prototypes()
    assert len(protos) >= 55
    for name, (argtypes, restype) in _lib._SIGS.items():
        assert name in protos, 'binding %s has no prototype in b200sp.h' % name
        params = protos[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes), params)
        for p, t in zip(params, argtypes):
            is_ptr_c = '*' in p
            is_ptr_py = t is C.c_void_p or hasattr(t, 'contents')
            assert is_ptr_c == is_ptr_py, (name, p, t)
            if not is_ptr_c:
                base = p.replace('const', '').split()[0]
                want = {'int': (C.c_int,), 'int32_t': (C.c_int,), 'float': (C.c_float,), 'double': (C.c_double,),
                        'int64_t': (C.c_int64,), 'uint64_t': (C.c_uint64,), 'uint32_t': (C.c_uint32,), 'size_t': (C.c_size_t, C.c_uint64)}[base]
                assert t in want, (name, p, t)


def test_struct_layouts_match_the_compiler(tmp_path):
    """sizeof/offsetof of every struct in b200sp.h as gcc lays it out == the ctypes mirrors in _lib.py."""
    import ctypes as C
    import subprocess
    from speedplusbaseline_b200 import _lib
    structs = {'b200sp_vtensor': _lib.VTensor, 'b200sp_bnfwd': _lib.BnFwd, 'b200sp_bnbwd': _lib.BnBwd,
               'b200sp_adamw_hp': _lib.AdamWHp, 'b200sp_aug': _lib.Aug, 'b200sp_convtc_chunk': _lib.ConvChunk,
               'b200sp_convtc_desc': _lib.ConvDesc, 'b200sp_in_apply_desc': _lib.InApplyDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200sp.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines += ['return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = str(tmp_path / 'layout')
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', exe], check=True, capture_output=True)
    got = dict(l.split() for l in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got['%s.%s' % (cname, f)]) == getattr(cls, f).offset, (cname, f)
