"""GPU bring-up script (not a pytest): per-layer forward and per-key gradient comparison of the
CUDA KRN path against the oracle.  Run on the B200 box:  python tests/gpu_debug_krn.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import krn as okrn, synth, steps  # noqa: E402
from speedplusbaseline_b200 import _lib as L  # noqa: E402
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet  # noqa: E402
from speedplusbaseline_b200.optim import FusedAdamW  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.manual_seed(0)
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    x, y = synth.synth_images(B), synth.synth_keypoints(B)
    m = KeypointRegressionNet(11)
    m.load_state_dict(sd)
    # round trip
    sd2 = m.state_dict()
    bad = [k for k in sd if not torch.equal(sd[k].cpu(), sd2[k].cpu())]
    print('state_dict round-trip mismatches:', bad[:5], 'order ok:', list(sd2) == list(sd))

    for train in (False, True):
        sdo = {k: v.clone() for k, v in sd.items()}
        m.load_state_dict(sd)
        m.train(train)
        taps = {}
        with torch.no_grad():
            feat, logits = okrn.krn_logits(sdo, x, train, taps=taps)
        cx = m.engine.forward(x.cuda(), y.cuda(), train=train)
        torch.cuda.synchronize()
        print('==== forward train=%s' % train)
        names = {'base.0.0': 'stem'}
        for b in m.engine.blocks:
            i, j = b['idx'], (0 if b['t'] == 1 else 1)
            if b['t'] != 1:
                names['base.%d.conv.0.0' % i] = 'e%d' % i
            names['base.%d.conv.%d.0' % (i, j)] = 'd%d' % i
            names['base.%d.conv.%d' % (i, j + 1)] = 'p%d' % i
        for e in (0, 1, 3):
            names['extras.%d.conv.0' % e] = 'xd%d' % e
            names['extras.%d.conv.3' % e] = 'xp%d' % e
        names['extras.2.conv.0'] = 'xr'
        for k, n in names.items():
            r, mx = rel(nhwc_to_nchw(cx.Y[n]), taps[k])
            flag = '' if r < 1e-4 else '   <<<<<<'
            print('  %-22s %-6s rel %.3e max %.3e%s' % (k, n, r, mx, flag))
        for i in (1, 2, 3, 13, 17):
            r, mx = rel(nhwc_to_nchw(cx.O[i]), taps['base.%d' % i])
            print('  block out %-3d rel %.3e max %.3e' % (i, r, mx))
        r, mx = rel(nhwc_to_nchw(cx.O['cat']), taps['extras.2'])
        print('  cat           rel %.3e max %.3e' % (r, mx))
        r, mx = rel(cx.logits, logits)
        print('  logits        rel %.3e max %.3e   ref range %.3e' % (r, mx, float(logits.abs().max())))
        if train:
            sdm = m.state_dict()
            worst = max((rel(sdm[k], sdo[k])[0], k) for k in sd if 'running' in k)
            print('  running stats worst rel', worst)

    # ---- one full train step: grads, clip, AdamW
    print('==== train step')
    sdo = {k: v.clone() for k, v in sd.items()}
    st = steps.new_state(sdo)
    ref = steps.krn_train_step(sdo, st, x, y)
    m.load_state_dict(sd)
    m.train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, clip_mode=1)
    opt.zero_grad()
    loss, sm = m(x.cuda(), y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print('  loss %.6f ref %.6f  | loss_x %.6f/%.6f loss_y %.6f/%.6f' % (float(loss), ref['loss'], sm['loss_x'], ref['loss_x'],
                                                                     sm['loss_y'], ref['loss_y']))
    gd = m.grad_dict()
    rows = []
    for k, g in ref['grads'].items():
        r, mx = rel(gd[k], g)
        rows.append((r, k, float(g.norm())))
    rows.sort(reverse=True)
    print('  worst grads (rel L2, key, ref norm):')
    for r in rows[:12]:
        print('    %.3e  %-40s %.3e' % r)
    print('  median grad rel %.3e' % rows[len(rows) // 2][0])
    opt.step()
    torch.cuda.synchronize()
    print('  grad norm %.6f ref %.6f' % (opt.last_grad_norm(), ref['grad_norm']))
    sdm = m.state_dict()
    rows = sorted(((rel(sdm[k], sdo[k])[0], k) for k in sd if okrn.is_param(k)), reverse=True)
    print('  worst post-step params:', rows[:5])
    print('  launches:', L.lib.b200sp_launch_count())


if __name__ == '__main__':
    main()
