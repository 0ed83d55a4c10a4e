"""Python emulation of dwr_bwd2_kernel's per-item index arithmetic (vectors over channels), checked against torch autograd."""
import numpy as np, torch, torch.nn.functional as F

def emu(S, B, H, W, C, g, y2, cA, cB, cC, yin, sc, shf, w9, SEG, relu6=True):
    Ho, Wo = (H - 1)//S + 1, (W - 1)//S + 1
    gq, yq, yb = g.reshape(-1), y2.reshape(-1), yin.reshape(-1)      # flat [B,Ho,Wo,C] / [B,H,W,C]
    g_in = np.full(B*H*W*C, np.nan, np.float64)
    dw = np.zeros((9, C)); ls1 = np.zeros(C); ls2 = np.zeros(C)
    cv = np.arange(C)
    def ld(arr, off): return arr[off + cv]
    def dyv(off, m): return (cA*ld(gq, off) + (cB*ld(yq, off) + cC)) * m
    actf = (lambda z: np.minimum(np.maximum(z, 0), 6)) if relu6 else (lambda z: np.maximum(z, 0))
    actd = (lambda z: (z > 0) & (z < 6)) if relu6 else (lambda z: z > 0)
    def finish(dg, yi, z):
        nonlocal ls1, ls2
        dg = np.where(actd(z), dg, 0.0); ls1 = ls1 + dg; ls2 = ls2 + dg*yi; return dg
    rows = H if S == 1 else Ho
    ncg = (W + 1)//2 if S == 1 else Wo
    nseg = (rows + SEG - 1)//SEG
    for b in range(B):
      for seg in range(nseg):
        r_a, r_b = seg*SEG, min(rows, seg*SEG + SEG)
        for wg in range(ncg):
          if S == 2:
            cb = wg; c1 = cb + 1 < Wo; w1 = 2*cb + 1 < W; m1 = 1.0 if c1 else 0.0
            dcol = C if c1 else 0; drow = Wo*C; xrow = W*C; xcol = C if w1 else 0
            od = ((b*Ho + r_a)*Wo + cb)*C; ox = ((b*H + 2*r_a)*W + 2*cb)*C
            E00, E01 = dyv(od, 1.0), dyv(od + dcol, m1)
            for a in range(r_a, r_b):
                r1 = a + 1 < Ho; h1 = 2*a + 1 < H
                odn = od + (drow if r1 else 0); mr = 1.0 if r1 else 0.0
                E10 = dyv(odn, mr); E11 = dyv(odn + dcol, mr*m1)
                oxh = ox + (xrow if h1 else 0)
                yi0, yi1, yi2, yi3 = ld(yb, ox), ld(yb, ox + xcol), ld(yb, oxh), ld(yb, oxh + xcol)
                z = yi0*sc + shf; av = actf(z); dg = E00*w9[4]; dw[4] += av*E00; g_in[ox + cv] = finish(dg, yi0, z)
                if w1:
                    z = yi1*sc + shf; av = actf(z); dg = E00*w9[5] + E01*w9[3]; dw[5] += av*E00; dw[3] += av*E01
                    g_in[ox + xcol + cv] = finish(dg, yi1, z)
                if h1:
                    z = yi2*sc + shf; av = actf(z); dg = E00*w9[7] + E10*w9[1]; dw[7] += av*E00; dw[1] += av*E10
                    g_in[oxh + cv] = finish(dg, yi2, z)
                if h1 and w1:
                    z = yi3*sc + shf; av = actf(z); dg = E00*w9[8] + E01*w9[6] + E10*w9[2] + E11*w9[0]
                    dw[8] += av*E00; dw[6] += av*E01; dw[2] += av*E10; dw[0] += av*E11
                    g_in[oxh + xcol + cv] = finish(dg, yi3, z)
                E00, E01 = E10, E11; od += drow; ox += 2*xrow
          else:
            wi0 = wg*2; p1 = wi0 + 1 < W
            dco = [min(max(wi0 - 1 + j, 0), Wo - 1)*C for j in range(4)]
            mc = [1.0 if 0 <= wi0 - 1 + j < Wo else 0.0 for j in range(4)]
            drow = Wo*C; dbase = b*Ho*Wo*C
            def dyrow(row):
                mr = 1.0 if 0 <= row < Ho else 0.0; o = dbase + min(max(row, 0), Ho - 1)*drow
                return [dyv(o + dco[j], mr*mc[j]) for j in range(4)]
            D0, D1 = dyrow(r_a - 1), dyrow(r_a)
            ox = ((b*H + r_a)*W + wi0)*C; xcol = C if p1 else 0; xrow = W*C
            for hi in range(r_a, r_b):
                rn = hi + 1; mr = 1.0 if rn < Ho else 0.0; o = dbase + min(rn, Ho - 1)*drow
                D2 = [dyv(o + dco[j], mr*mc[j]) for j in range(4)]
                for o2 in range(2):
                    if o2 == 0 or p1:
                        yi = ld(yb, ox + (xcol if o2 else 0)); z = yi*sc + shf; av = actf(z); dg = np.zeros(C)
                        for sx in range(3):
                            dg = dg + D0[o2+sx]*w9[6+(2-sx)] + D1[o2+sx]*w9[3+(2-sx)] + D2[o2+sx]*w9[0+(2-sx)]
                            dw[6+(2-sx)] += av*D0[o2+sx]; dw[3+(2-sx)] += av*D1[o2+sx]; dw[0+(2-sx)] += av*D2[o2+sx]
                        g_in[ox + (xcol if o2 else 0) + cv] = finish(dg, yi, z)
                D0, D1 = D1, D2; ox += xrow
    return g_in.reshape(B, H, W, C), dw, ls1, ls2

def ref(S, B, H, W, C, g, y2, cA, cB, cC, yin, sc, shf, w9, relu6=True):
    z = torch.tensor(yin*sc + shf)                                   # [B,H,W,C]
    x = (z.clamp(0, 6) if relu6 else z.clamp(min=0)).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w = torch.tensor(w9.reshape(3, 3, C)).permute(2, 0, 1).reshape(C, 1, 3, 3).contiguous().requires_grad_(True)
    out = F.conv2d(x, w, None, S, 1, 1, C)
    dy = torch.tensor(cA*g + cB*y2 + cC).permute(0, 3, 1, 2)
    out.backward(dy)
    d = ((z > 0) & (z < 6)) if relu6 else (z > 0)
    gi = x.grad.permute(0, 2, 3, 1) * d
    dw = w.grad.reshape(C, 9).t()
    return gi.numpy(), dw.numpy(), gi.sum((0, 1, 2)).numpy(), (gi*torch.tensor(yin)).sum((0, 1, 2)).numpy()

rng = np.random.default_rng(0)
for S in (1, 2):
    for (B, H, W, C, SEG, r6) in [(2, 6, 6, 4, 2, True), (1, 7, 5, 8, 3, True), (2, 8, 12, 4, 8, False), (1, 5, 7, 4, 2, True), (1, 14, 14, 4, 8, True)]:
        Ho, Wo = (H-1)//S + 1, (W-1)//S + 1
        g, y2 = rng.normal(size=(B, Ho, Wo, C)), rng.normal(size=(B, Ho, Wo, C))
        cA, cB, cC = rng.normal(size=C), rng.normal(size=C)*0.3, rng.normal(size=C)*0.1
        yin = rng.normal(size=(B, H, W, C))*2; sc, shf = rng.uniform(0.5, 2, C), rng.normal(size=C) + 1.5
        w9 = rng.normal(size=(9, C))
        a = emu(S, B, H, W, C, g, y2, cA, cB, cC, yin, sc, shf, w9, SEG, r6)
        r = ref(S, B, H, W, C, g, y2, cA, cB, cC, yin, sc, shf, w9, r6)
        errs = [float(np.abs(x - y).max()) for x, y in zip(a, r)]
        print('S=%d %s' % (S, (B, H, W, C, SEG, r6)), 'nan' if np.isnan(a[0]).any() else 'ok', ['%.1e' % e for e in errs])
