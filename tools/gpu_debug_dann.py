"""debug helper (not a test): per-tensor DANN gradient errors vs the float64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import revgrad as orev, synth, steps
from kutil import rel
from test_dann_gpu import _model, _inputs, _oracle, _stepper

B, alpha = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 0.37
sd = synth.synth_state_dict(orev.revgrad_shapes(), 2021)
src, lab, tgt = _inputs(B)
r64, s64 = _oracle(sd, src, lab, tgt, alpha, torch.float64)
r32, s32 = _oracle(sd, src, lab, tgt, alpha, torch.float32)
m = _model(sd).train()
stp, opt = _stepper(m, False)
stp._set_alpha(alpha)
losses = stp._fwd_bwd(src.cuda(), lab.cuda(), tgt.cuda()).cpu()
torch.cuda.synchronize()
print('losses', losses.tolist(), [r64['loss_pose'], r64['loss_source'], r64['loss_target']])
gd = m.grad_dict()
gn = r64['grad_norm']
print('grad_norm64', gn, 'f32', r32['grad_norm'])
rows = []
for k, g64 in r64['grads'].items():
    if float(g64.norm()) < 1e-3 * gn:
        continue
    e_cuda, e_f32 = rel(gd[k], g64), rel(r32['grads'][k], g64)
    rows.append((e_cuda / (e_f32 + 1e-4), k, e_cuda, e_f32, float(g64.norm())))
rows.sort(reverse=True)
for r in rows[:12]:
    print('%8.2f %-40s cuda %.3e f32 %.3e |g| %.3e' % r)
print('mean ratio', sum(r[0] for r in rows) / len(rows))
for k in ('domain_classifier.0.weight', 'domain_classifier.0.bias', 'domain_classifier.3.weight', 'domain_classifier.3.bias'):
    print(k, rel(gd[k], r64['grads'][k]))
