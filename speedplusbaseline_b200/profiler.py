"""Per-launch CUDA-event profiler for the eager (un-graphed) step: wraps _lib.call, brackets every
libb200sp launch with events on the launching stream and attributes ALGORITHMIC bytes (each distinct
operand tensor read or written once; DESIGN.md "Kernels") to it.  Used by bench.py for the `roofline`
object and by profiles/*.txt."""
import json
import os
from collections import OrderedDict, defaultdict

import torch

from . import _lib as L


def _vt(a):
    return a._obj if hasattr(a, '_obj') else None


def algorithmic_bytes(name, a, elt=4):
    """bytes one launch must move (operands once each); a = the ctypes argument tuple."""
    try:
        if name == 'b200sp_pw_fwd':
            M, N, K = a[6], a[7], a[8]
            return (M * K + M * N) * elt + N * K * 4
        if name == 'b200sp_pw_dgrad':
            M, N, K = a[6], a[7], a[8]
            two = 2 if _vt(a[0]).mode == L.VT_DY else 1
            b = two * M * N * elt + N * K * 4 + M * K * elt
            if a[5] is not None:
                b += M * K * elt
            if a[2] is not None:
                b += M * K * elt
            return b
        if name == 'b200sp_pw_wgrad':
            M, N, K = a[4], a[5], a[6]
            two = 2 if _vt(a[0]).mode == L.VT_DY else 1
            return two * M * N * elt + M * K * elt + 2 * N * K * 4
        if name == 'b200sp_dw_fwd':
            B, H, W, C, s = a[4], a[5], a[6], a[7], a[8]
            Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
            return B * C * (H * W + Ho * Wo) * elt
        if name == 'b200sp_dw_bwd':
            B, H, W, C, s = a[7], a[8], a[9], a[10], a[11]
            Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
            two = 2 if _vt(a[0]).mode == L.VT_DY else 1
            b = B * C * (two * Ho * Wo + 2 * H * W) * elt       # dy(+y), input value, g_in
            if a[3] is not None:
                b += B * C * H * W * elt
            if a[6] is not None and _vt(a[6]) is not None and _vt(a[6]).y != _vt(a[1]).x:
                b += B * C * H * W * elt
            return b
        if name == 'b200sp_stem_fwd':
            B, H, W = a[4], a[5], a[6]
            return B * 3 * H * W * 4 + B * (H // 2) * (W // 2) * 32 * elt
        if name == 'b200sp_stem_wgrad':
            B, H, W = a[3], a[4], a[5]
            return B * 3 * H * W * 4 + 2 * B * (H // 2) * (W // 2) * 32 * elt
        if name == 'b200sp_bn_apply':
            M, C = a[6], a[7]
            return M * C * elt * (2 + (1 if a[3] is not None else 0))
        if name == 'b200sp_bn_bwd_reduce':
            return 2 * a[2] * a[3] * elt
        if name == 'b200sp_reorg_cat_fwd':
            B, h, w, Cr, C1 = a[3], a[4], a[5], a[6], a[7]
            return 2 * B * h * w * (4 * Cr + C1) * elt
        if name == 'b200sp_reorg_cat_bwd':
            B, h, w, Cr, C1 = a[5], a[6], a[7], a[8], a[9]
            return 5 * B * h * w * (4 * Cr + C1) * elt
        if name == 'b200sp_head_fwd':
            B, HWC, N = a[3], a[4], a[6]
            return B * HWC * elt + N * HWC * 4
        if name == 'b200sp_head_bwd':
            B, HWC, N = a[7], a[8], a[10]
            return 2 * B * HWC * elt + 3 * N * HWC * 4
        if name == 'b200sp_adamw_step':
            return 28 * a[5]
        if name == 'b200sp_grad_sqnorm':
            return 4 * a[1]
    except Exception:
        return 0
    return 0


class LaunchTimer:
    def __init__(self):
        self.records = []          # (name, bytes, ev0, ev1, shape-tag)
        self._orig = None

    def __enter__(self):
        self._orig = L.call

        def timed(name, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._orig(name, *args)
            e1.record()
            tag = ','.join(str(x) for x in args if isinstance(x, int) and not isinstance(x, bool) and 0 < x < (1 << 24))
            self.records.append((name, algorithmic_bytes(name, args), e0, e1, tag))
        L.call = timed
        # engine modules bound `L` as a module alias, so patching the attribute is enough
        return self

    def __exit__(self, *exc):
        L.call = self._orig
        torch.cuda.synchronize()
        return False

    def rows(self):
        return [(n, b, e0.elapsed_time(e1) * 1e3, tag) for n, b, e0, e1, tag in self.records]   # us


def profile_krn_step(stepper, images, target, reps=3):
    best = None
    # per-kernel timing needs kernels run ONE AT A TIME: switch off the side-stream overlap of the weight-gradient
    # GEMMs for the profiled steps (it is restored afterwards; the timed bench loop keeps it on)
    eng = stepper.model.engine
    saved = getattr(eng, '_async_wgrad', False)
    eng._async_wgrad = False
    try:
        best = _profile_reps(stepper, images, target, reps)
    finally:
        eng._async_wgrad = saved
    return _summarise(best, reps)


def _profile_reps(stepper, images, target, reps):
    best = None
    for _ in range(reps):
        with LaunchTimer() as lt:
            # keep the GPU busy for ~8 ms first: the host then runs ahead and every (event, launch, event) triple of the step is
            # already queued when the GPU reaches it -- the intervals are kernel time, not kernel time + host launch latency
            # (small kernels read 5-10 us too long otherwise)
            torch.cuda._sleep(16_000_000)
            stepper.eager(images, target)
        rows = lt.rows()
        if best is None:
            best = rows
        else:
            best = [(n, b, min(t, t2), tag) for (n, b, t, tag), (_, _, t2, _) in zip(best, rows)]
    return best


def _summarise(best, reps):
    agg = defaultdict(lambda: [0, 0.0, 0])
    for n, b, t, _ in best:
        agg[n][0] += 1
        agg[n][1] += t
        agg[n][2] += b
    total = sum(v[1] for v in agg.values())
    step_bytes = sum(v[2] for v in agg.values())
    lines = ['# per-kernel-family totals for ONE eager KRN train step (CUDA events around every launch, launches queued ahead of the GPU, min of %d reps)' % reps,
             '%-24s %6s %10s %7s %12s %9s' % ('kernel', 'calls', 'us', 'share', 'alg_MB', 'GB/s')]
    fam = sorted(agg.items(), key=lambda kv: -kv[1][1])
    for n, (c, t, b) in fam:
        lines.append('%-24s %6d %10.1f %6.1f%% %12.2f %9.1f' % (n, c, t, 100 * t / total, b / 1e6, b / 1e3 / max(t, 1e-3)))
    lines.append('%-24s %6d %10.1f %6.1f%% %12.2f %9.1f' % ('TOTAL', len(best), total, 100.0, step_bytes / 1e6, step_bytes / 1e3 / total))
    lines.append('')
    lines.append('# top 25 individual launches')
    top = sorted(best, key=lambda r: -r[2])[:25]
    for n, b, t, tag in top:
        lines.append('%-24s %10.1f us %10.2f MB %9.1f GB/s   [%s]' % (n, t, b / 1e6, b / 1e3 / max(t, 1e-3), tag))
    lines.append('')
    lines.append('# every launch in issue order')
    for n, b, t, tag in best:
        lines.append('%-24s %10.1f us %10.2f MB %9.1f GB/s   [%s]' % (n, t, b / 1e6, b / 1e3 / max(t, 1e-3), tag))
    n, b, t, tag = top[0]
    hbm = 6545.9
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        hbm = json.load(open(p)).get('hbm_gbs', hbm)
    # kernel FAMILIES: the three pointwise-GEMM entry points are one kernel family (tcgemm2.cu / tcgemm.cu), the two depthwise
    # entry points another (dwroll.cu); the roofline object of the bench line leads with the family that owns the step
    groups = OrderedDict((('pointwise_gemm', ('b200sp_pw_fwd', 'b200sp_pw_dgrad', 'b200sp_pw_wgrad')),
                          ('depthwise', ('b200sp_dw_fwd', 'b200sp_dw_bwd')),
                          ('stem', ('b200sp_stem_fwd', 'b200sp_stem_wgrad')),
                          ('optimizer', ('b200sp_adamw_step', 'b200sp_grad_sqnorm', 'b200sp_optim_step'))))
    fams, seen = [], set()
    for gname, members in groups.items():
        c = sum(agg[m][0] for m in members if m in agg)
        t_ = sum(agg[m][1] for m in members if m in agg)
        b_ = sum(agg[m][2] for m in members if m in agg)
        seen.update(members)
        if c:
            fams.append({'family': gname, 'launches': c, 'us': t_, 'share': t_ / total, 'algorithmic_mb': b_ / 1e6,
                         'gbs': b_ / 1e3 / max(t_, 1e-3), 'frac_of_hbm_peak': b_ / 1e3 / max(t_, 1e-3) / hbm})
    c = sum(v[0] for k, v in agg.items() if k not in seen)
    t_ = sum(v[1] for k, v in agg.items() if k not in seen)
    b_ = sum(v[2] for k, v in agg.items() if k not in seen)
    fams.append({'family': 'other (bn_apply, head, reorg, loss)', 'launches': c, 'us': t_, 'share': t_ / total, 'algorithmic_mb': b_ / 1e6,
                 'gbs': b_ / 1e3 / max(t_, 1e-3), 'frac_of_hbm_peak': b_ / 1e3 / max(t_, 1e-3) / hbm})
    fams.sort(key=lambda f: -f['us'])
    return {'table': '\n'.join(lines) + '\n', 'step_bytes': step_bytes, 'step_roofline_ms': step_bytes / hbm / 1e6,
            'eager_kernel_ms': total / 1e3, 'families': fams,
            'dominant': {'name': '%s[%s]' % (n, tag), 'us': t, 'bytes': b, 'gbs': b / 1e3 / max(t, 1e-3),
                         'share': t / total}}
