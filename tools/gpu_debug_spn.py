import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import spn as ospn, synth
from kutil import rel
from speedplusbaseline_b200.nets.spn import SpacecraftPoseNet
from speedplusbaseline_b200.optim import FusedAdamW
from speedplusbaseline_b200.core.trainer import SPNTrainStep
sd = synth.synth_state_dict(ospn.spn_shapes(), 2021)
B = 4
x = synth.synth_images(B, 227, 227, seed=9)
yc, yw = synth.synth_soft_targets(B, tag='cls'), synth.synth_soft_targets(B, tag='wts')
m = SpacecraftPoseNet(5000, pretrain=False, device='cuda:0'); m.load_state_dict(sd); m.train(); m.engine.drop_p = 0.0
opt = FusedAdamW(m._store, m.parameters(), clip_mode=2)
stp = SPNTrainStep(m, opt, use_graph=False)
stp._fwd_bwd(x.cuda(), yc.cuda(), yw.cuda()); torch.cuda.synchronize()
bf = m.engine._bufs
W7, W8, W6 = sd['fc7.weight'].cuda().double(), sd['fc8.weight'].cuda().double(), sd['fc6.weight'].cuda().double()
dz, h2, h1, f = bf['dz_c'].double(), bf['h_fc7'].double(), bf['h_fc6'].double(), bf['f'].double()
dh2_ref = (dz @ W8) * (h2 > 0)
print('dh2', rel(bf['dh_fc7'], dh2_ref))
dh1_ref = (bf['dh_fc7'].double() @ W7) * (h1 > 0)
print('dh1', rel(bf['dh_fc6'], dh1_ref), 'nnz mask', float((h1 > 0).double().mean()))
d = (bf['dh_fc6'].double() - dh1_ref).abs()
print('max abs err', float(d.max()), 'ref max', float(dh1_ref.abs().max()), 'rows err', d.sum(1).tolist())
cols = d.sum(0); print('worst cols', torch.topk(cols, 8))
nomask = bf['dh_fc7'].double() @ W7
print('vs unmasked', rel(bf['dh_fc6'], nomask))
# ---- compare against the float64 oracle
with torch.no_grad():
    sd64 = {k: v.double() for k, v in sd.items()}
    f64 = ospn.spn_features(sd64, x.double())
z6 = f64 @ sd64['fc6.weight'].t() + sd64['fc6.bias']
ours_f = bf['f'].permute(0, 3, 1, 2).reshape(B, -1).double().cpu()
print('f rel err', rel(ours_f, f64), 'mean ratio ours/ref', float((ours_f[f64 > 0.1] / f64[f64 > 0.1]).mean()))
m_ref, m_our = (z6 > 0), (bf['h_fc6'].cpu() > 0)
print('mask mismatches', int((m_ref != m_our).sum()), 'of', m_ref.numel())
h1_ref = torch.relu(z6)
print('h1 rel err', rel(bf['h_fc6'], h1_ref))
bad = (m_ref != m_our).nonzero()[:10]
for b, j in bad.tolist():
    print('  z6 ref %.3e ours %.3e' % (float(z6[b, j]), float(bf['h_fc6'][b, j])))
# ---- oracle dL/dz6 via autograd
import torch.nn.functional as F
f_ = f64.clone().requires_grad_(False)
z6v = (f_ @ sd64['fc6.weight'].t() + sd64['fc6.bias']).requires_grad_(True)
h = F.relu(z6v)
h2v = F.relu(h @ sd64['fc7.weight'].t() + sd64['fc7.bias'])
c = h2v @ sd64['fc8.weight'].t() + sd64['fc8.bias']
loss = ospn.soft_ce(c, yc.double())
loss.backward()
g_ref = z6v.grad
ours = bf['dh_fc6'].double().cpu()
print('dL/dz6 rel', rel(ours, g_ref), 'rows', [rel(ours[i], g_ref[i]) for i in range(B)])
print('colsum rel', rel(ours.sum(0), g_ref.sum(0)))
gb = m.grad_dict()['fc6.bias'].double().cpu()
print('our fc6.bias vs colsum(ours)', rel(gb, ours.sum(0)), ' vs oracle colsum', rel(gb, g_ref.sum(0)))
