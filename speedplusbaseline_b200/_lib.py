"""ctypes binding of libb200sp.so (include/b200sp.h).  This is the reference-side stub of
INTEGRATION.md.  There is NO fallback: if the library is missing the import fails loudly."""
import ctypes as C
import os

import torch

from . import _build

F32, BF16, F32_TF32X1 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_RELU6, ACT_LEAKY02, ACT_SIGMOID = 0, 1, 2, 3, 4
VT_PLAIN, VT_BNACT, VT_DY = 0, 1, 2
OPT_ADAMW, OPT_SGD, OPT_RMSPROP, OPT_ADAM = 0, 1, 2, 3
BF16_ENABLED = os.environ.get('B200SP_ENABLE_BF16', '0') == '1'   # experimental bf16 storage for --use_fp16

vp = C.c_void_p


class VTensor(C.Structure):
    _fields_ = [('x', vp), ('x2', vp), ('p0', vp), ('p1', vp), ('p2', vp), ('mode', C.c_int32), ('act', C.c_int32)]


class BnFwd(C.Structure):
    _fields_ = [('sum', vp), ('sumsq', vp), ('ticket', vp), ('gamma', vp), ('beta', vp),
                ('running_mean', vp), ('running_var', vp), ('scale', vp), ('shift', vp),
                ('mean', vp), ('rstd', vp), ('momentum', C.c_float), ('eps', C.c_float)]


class BnBwd(C.Structure):
    _fields_ = [('s1', vp), ('s2', vp), ('ticket', vp), ('y', vp), ('scale', vp), ('shift', vp),
                ('mean', vp), ('rstd', vp), ('cA', vp), ('cB', vp), ('cC', vp), ('dgamma', vp), ('dbeta', vp),
                ('act', C.c_int32), ('pad_', C.c_int32)]


class AdamWHp(C.Structure):
    _fields_ = [('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
                ('weight_decay', C.c_float), ('max_norm', C.c_float), ('clip_value', C.c_float),
                ('grad_scale', C.c_float), ('step', C.c_int32), ('clip_mode', C.c_int32),
                ('sqnorm', C.c_double), ('last_norm', C.c_float), ('pad_', C.c_float)]


class Aug(C.Structure):          # b200sp_aug (include/b200sp.h): per-image crop box + augmentation decisions
    _fields_ = [('x0', C.c_int32), ('x1', C.c_int32), ('y0', C.c_int32), ('y1', C.c_int32), ('rot', C.c_int32),
                ('flip', C.c_int32), ('bc', C.c_int32), ('a', C.c_float), ('b', C.c_float), ('noise_std', C.c_float),
                ('seed', C.c_uint32), ('pad_', C.c_int32)]


class ConvChunk(C.Structure):
    _fields_ = [('plane', C.c_int32), ('c0', C.c_int32), ('shift', C.c_int32), ('pad_', C.c_int32)]


CONVTC_MAX_CHUNKS = 64


class ConvDesc(C.Structure):
    _fields_ = [('planes', vp * 4), ('w', vp), ('out', vp), ('stats', vp),
                ('C', C.c_int32), ('B', C.c_int32), ('Hq', C.c_int32), ('Wq', C.c_int32), ('Ho', C.c_int32), ('Wo', C.c_int32),
                ('N_pad', C.c_int32), ('N_out', C.c_int32), ('n_chunks', C.c_int32), ('pad_', C.c_int32),
                ('chunks', ConvChunk * CONVTC_MAX_CHUNKS),
                ('OH', C.c_int32), ('OW', C.c_int32), ('sy', C.c_int32), ('sx', C.c_int32), ('oy', C.c_int32), ('ox', C.c_int32)]


class InApplyDesc(C.Structure):
    _fields_ = [('raw', vp), ('scale', vp), ('shift', vp), ('res_in', vp), ('res_out', vp), ('planes', vp * 4),
                ('B', C.c_int32), ('Hs', C.c_int32), ('Ws', C.c_int32), ('Cs', C.c_int32), ('C', C.c_int32), ('act', C.c_int32),
                ('pad', C.c_int32), ('up', C.c_int32), ('ps', C.c_int32), ('Hd', C.c_int32), ('Wd', C.c_int32), ('Cd', C.c_int32),
                ('pad_mode', C.c_int32), ('pad2_', C.c_int32)]


def _load():
    path = _build.LIB
    if not os.path.exists(path):
        if os.environ.get('B200SP_NO_AUTOBUILD'):
            raise ImportError('libb200sp.so missing (%s); run `python -c "import __graft_entry__ as g; g.build()"`' % path)
        _build.build()
    return C.CDLL(path)


lib = _load()
i32, i64, f32, f64 = C.c_int, C.c_int64, C.c_float, C.c_double
PVT, PBF, PBB = C.POINTER(VTensor), C.POINTER(BnFwd), C.POINTER(BnBwd)

_SIGS = {
    'b200sp_version': ([], i32),
    'b200sp_launch_count': ([], i64),
    'b200sp_set_workspace': ([vp, C.c_size_t], i32),
    'b200sp_tc_probe': ([vp, vp, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_mma_probe': ([i32, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    'b200sp_stem_fwd': ([vp, vp, vp, PBF, i32, i32, i32, i32, vp], i32),
    'b200sp_stem_wgrad': ([vp, PVT, vp, i32, i32, i32, i32, vp], i32),
    'b200sp_pw_fwd': ([PVT, vp, vp, i32, vp, PBF, i32, i32, i32, i32, vp], i32),
    'b200sp_pw_dgrad': ([PVT, vp, vp, f32, vp, PBB, i32, i32, i32, i32, vp], i32),
    'b200sp_pw_wgrad': ([PVT, PVT, vp, vp, i32, i32, i32, i32, vp], i32),
    'b200sp_dw_fwd': ([PVT, vp, vp, PBF, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_dw_bwd': ([PVT, PVT, vp, vp, vp, vp, PBB, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_bn_apply': ([vp, vp, vp, vp, i32, vp, i64, i32, i32, vp], i32),
    'b200sp_bn_eval_affine': ([vp, vp, vp, vp, f32, vp, vp, i64, vp], i32),
    'b200sp_bn_fwd_finalize': ([PBF, i32, f64, vp], i32),
    'b200sp_bn_bwd_finalize': ([PBB, i32, f64, vp], i32),
    'b200sp_bn_bwd_reduce': ([vp, PBB, i64, i32, i32, vp], i32),
    'b200sp_add_i64': ([vp, i64, i64, vp], i32),
    'b200sp_reorg_cat_fwd': ([PVT, PVT, vp, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_reorg_cat_bwd': ([vp, vp, vp, PBB, PBB, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_head_fwd': ([PVT, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_head_bias': ([vp, vp, i32, i32, vp], i32),
    'b200sp_krn_loss': ([vp, vp, vp, vp, vp, vp, i32, i32, vp], i32),
    'b200sp_head_bwd': ([vp, PVT, vp, vp, vp, vp, PBB, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_convtc_fwd': ([C.POINTER(ConvDesc), vp], i32),
    'b200sp_conv_kwsum': ([vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_sa_prep': ([vp, vp, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_in_finalize': ([vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, f32, vp], i32),
    'b200sp_in_apply': ([C.POINTER(InApplyDesc), vp], i32),
    'b200sp_in_apply_final': ([vp, vp, vp, vp, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_style_embed': ([vp, vp, vp, vp, f32, vp, i32, i32, vp], i32),
    'b200sp_style_linear': ([vp, vp, vp, vp, i32, i32, i32, vp], i32),
    'b200sp_gemm_fwd': ([PVT, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_gemm_dgrad': ([PVT, i32, vp, vp, f32, vp, PBB, i32, i32, i32, i32, vp], i32),
    'b200sp_gemm_wgrad': ([PVT, i32, PVT, i32, vp, i32, i32, i32, i32, vp], i32),
    'b200sp_colsum_f32': ([PVT, vp, i32, i32, i32, vp], i32),
    'b200sp_fc_fwd_splitk': ([vp, vp, vp, i32, i32, i32, vp], i32),
    'b200sp_fc_dgrad_splitk': ([vp, vp, vp, i32, i32, i32, vp], i32),
    'b200sp_bias_act': ([vp, vp, i32, i32, i32, vp], i32),
    'b200sp_im2col': ([vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_col2im': ([vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    'b200sp_pool_lrn_fwd': ([vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp], i32),
    'b200sp_pool_lrn_bwd': ([vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, i32, vp], i32),
    'b200sp_dropout_fwd': ([vp, vp, vp, i64, f32, C.c_uint64, vp], i32),
    'b200sp_dropout_bwd': ([vp, vp, i64, f32, vp], i32),
    'b200sp_dropout_fwd_ctr': ([vp, vp, vp, i64, f32, C.c_uint64, vp, vp], i32),
    'b200sp_soft_ce': ([vp, vp, vp, vp, i32, i32, f32, vp], i32),
    'b200sp_soft_ce_mean': ([vp, vp, vp, i32, vp], i32),
    'b200sp_relu_mask': ([vp, vp, i64, vp], i32),
    'b200sp_dann_head_fwd': ([vp, vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    'b200sp_bce_logits': ([vp, f32, vp, vp, vp, i32, vp], i32),
    'b200sp_dann_head_bwd': ([vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    'b200sp_scale_dev': ([vp, i64, vp, f32, i32, vp], i32),
    'b200sp_grad_sqnorm': ([vp, i64, vp, vp], i32),
    'b200sp_adamw_step': ([vp, vp, vp, vp, vp, i64, vp, vp], i32),
    'b200sp_optim_step': ([i32, vp, vp, vp, vp, vp, i64, vp, vp], i32),
    'b200sp_topk_softmax': ([vp, vp, vp, vp, i32, i32, i32, vp], i32),
    'b200sp_kpt_denorm': ([vp, vp, vp, i32, i32, vp], i32),
    'b200sp_input_pipeline': ([vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, i32, i32, i32, vp, vp], i32),
    'b200sp_kpt_augment': ([vp, vp, vp, i32, i32, i32, vp], i32),
}
for _n, (_a, _r) in _SIGS.items():
    _f = getattr(lib, _n)
    _f.argtypes, _f.restype = _a, _r

EXPORTS = tuple(_SIGS)


class B200SPError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        raise B200SPError('libb200sp %s failed with code %d' % (what, rc))


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}


def ensure_workspace(device, nbytes=96 << 20):
    """Process-lifetime scratch of the presplit GEMM route (include/b200sp.h b200sp_set_workspace), one per device, never freed:
    the library keeps the raw pointer."""
    key = torch.device(device).index or 0
    if key not in _workspaces:
        _workspaces[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    ws = _workspaces[key]
    call('b200sp_set_workspace', ws.data_ptr(), ws.numel())
    return ws


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda():
    if not torch.cuda.is_available():
        raise B200SPError('a CUDA device (B200, sm_100a) is required: there is no CPU fallback in '
                          'speedplusbaseline_b200; use --no_cuda to run the unmodified torch path')
