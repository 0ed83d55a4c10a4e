import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import krn as okrn, synth
from kutil import rel
from speedplusbaseline_b200 import _lib as L
from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
B = 4
sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
x, y = synth.synth_images(B), synth.synth_keypoints(B)
sdo = {k: v.clone().double() if v.is_floating_point() else v.clone() for k, v in sd.items()}
taps = {}
with torch.no_grad():
    taps['logits'] = okrn.krn_logits(sdo, x.double(), True, taps=taps)[1]
for dt in (L.BF16,):
    m = KeypointRegressionNet(11, device='cuda:0', dtype=dt); m.load_state_dict(sd); m.train()
    cx = m.engine.forward(x.cuda(), y.cuda(), train=True)
    torch.cuda.synchronize()
    names = [('base.0.0', 'stem')]
    for b in m.engine.blocks:
        i, j = b['idx'], (0 if b['t'] == 1 else 1)
        if b['t'] != 1:
            names.append(('base.%d.conv.0.0' % i, 'e%d' % i))
        names.append(('base.%d.conv.%d.0' % (i, j), 'd%d' % i))
        names.append(('base.%d.conv.%d' % (i, j + 1), 'p%d' % i))
    for e in (0, 1):
        names += [('extras.%d.conv.0' % e, 'xd%d' % e), ('extras.%d.conv.3' % e, 'xp%d' % e)]
    names += [('extras.2.conv.0', 'xr'), ('extras.3.conv.0', 'xd3'), ('extras.3.conv.3', 'xp3')]
    for k, n in names:
        print('%-22s %-5s %.3e' % (k, n, rel(cx.Y[n].float().permute(0, 3, 1, 2), taps[k])))
    print('logits', rel(cx.logits, taps['logits']))
