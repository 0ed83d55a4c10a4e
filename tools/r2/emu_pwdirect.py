"""numpy emulation of pwf_kernel's tile / chunk / statistics index logic (branch r2-candidates) against the plain definition."""
import numpy as np
def emu(x, w, sc, sh, act, K, N, NC, M, grid, warps=8):
    y = np.full((M, N), np.nan); s_w = np.zeros(K*N)
    for i in range(N*K): s_w[(i % K)*N + i//K] = w.reshape(-1)[i]
    xf = (lambda v, k: v) if sc is None else (lambda v, k: act(v*sc[k:k+len(v)] + sh[k:k+len(v)]))
    NCH = N//NC; tot_s = np.zeros(N); tot_q = np.zeros(N)
    ntiles = (M + 63)//64
    for blk in range(grid):
        s_red = np.zeros((warps, 2, N))
        for warp in range(warps):
            rs = np.zeros((NCH, 32)); rq = np.zeros((NCH, 32))           # [chunk][lane]
            t = blk*warps + warp
            while t < ntiles:
                acc_all = np.zeros((32, 2, N)); vv = np.zeros((32, 2), bool); pp = np.zeros((32, 2), int)
                for lane in range(32):
                    p0, p1 = t*64 + lane, t*64 + lane + 32
                    vv[lane] = (p0 < M, p1 < M); pp[lane] = (p0, p1)
                    rows = [x[min(p0, M-1)], x[min(p1, M-1)]]
                    xr = [np.concatenate([xf(r[k:k+4], k) for k in range(0, K, 4)]) for r in rows]
                    for ch in range(NCH):
                        for px in range(2):
                            for j in range(NC):
                                acc_all[lane, px, ch*NC + j] = sum(xr[px][k]*s_w[k*N + ch*NC + j] for k in range(K))
                for ch in range(NCH):
                    T = np.zeros((32, NC + 1)); Q = np.zeros((32, NC + 1))
                    for lane in range(32):
                        a0, a1 = acc_all[lane, 0, ch*NC:(ch+1)*NC], acc_all[lane, 1, ch*NC:(ch+1)*NC]
                        if vv[lane, 0]: y[pp[lane, 0], ch*NC:(ch+1)*NC] = a0
                        if vv[lane, 1]: y[pp[lane, 1], ch*NC:(ch+1)*NC] = a1
                        m0, m1 = float(vv[lane, 0]), float(vv[lane, 1])
                        T[lane, :NC] = m0*a0 + m1*a1; Q[lane, :NC] = m0*a0*a0 + m1*a1*a1
                    for lane in range(NC):
                        rs[ch, lane] += T[:, lane].sum(); rq[ch, lane] += Q[:, lane].sum()
                t += grid*warps
            for c in range(NCH):
                for lane in range(NC):
                    s_red[warp, 0, c*NC + lane] = rs[c, lane]; s_red[warp, 1, c*NC + lane] = rq[c, lane]
        tot_s += s_red[:, 0].sum(0); tot_q += s_red[:, 1].sum(0)
    return y, tot_s, tot_q
rng = np.random.default_rng(0)
relu6 = lambda z: np.minimum(np.maximum(z, 0), 6)
for (K, N, NC, M, grid, bn) in [(32, 16, 16, 200, 2, True), (16, 96, 32, 130, 1, True), (96, 24, 24, 70, 3, False), (24, 144, 24, 65, 1, True), (192, 32, 32, 64, 2, True)]:
    x = rng.normal(size=(M, K)); w = rng.normal(size=(N, K))
    sc, sh = (rng.uniform(.5, 2, K), rng.normal(size=K)) if bn else (None, None)
    y, s, q = emu(x, w, sc, sh, relu6, K, N, NC, M, grid)
    ref = (relu6(x*sc + sh) if bn else x) @ w.T
    print((K, N, NC, M, grid, bn), 'nan' if np.isnan(y).any() else 'ok', '%.1e %.1e %.1e' % (np.abs(y-ref).max(), np.abs(s-ref.sum(0)).max(), np.abs(q-(ref**2).sum(0)).max()))
