// dwconv.cu -- depthwise 3x3 convolution (forward; fused dgrad+wgrad+BN/activation backward) and
// the 3->32 stride-2 stem.  HBM-bound CUDA-core stencils over NHWC: every thread owns 4 channels
// (one 16-byte vector) and slides a 3x3 register window along a row segment, so each input
// element is fetched from L1/L2 three times instead of nine and from HBM once.
// Replaces torch conv2d(groups=C) at park2019.py:47 and torchvision mobilenetv2.py:45-49,126.
#include "common.cuh"

namespace {

constexpr int DW_NT = 256;

struct DwGeom {
    int B, H, W, C, Ho, Wo;
    int CB;            // channel-vectors (of 4) per CTA row; divides C/4
    int SPC;           // strips per CTA pass = DW_NT / CB
    int SEGW, nseg;    // row segment length / segments per row
    long long nstrips;
    double count;
};

__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4act(float4 z, ActP act) {
    return make_float4(act_fwd(z.x, act), act_fwd(z.y, act), act_fwd(z.z, act), act_fwd(z.w, act));
}
__device__ __forceinline__ float4 f4actbwd(float4 z, ActP act) {
    return make_float4(act_bwd(z.x, act), act_bwd(z.y, act), act_bwd(z.z, act), act_bwd(z.w, act));
}

// per-thread copy of a virtual tensor's channel parameters (4 channels)
struct VtParams {
    float4 p0, p1, p2;
    int mode;
    ActP act;
};
__device__ __forceinline__ VtParams vt_params(const b200sp_vtensor& t, int c) {
    VtParams p;
    p.mode = t.mode; p.act = act_params(t.act);
    p.p0 = p.p1 = p.p2 = f4zero();
    if (t.mode != B200SP_VT_PLAIN) { p.p0 = ldg4(t.p0 + c); p.p1 = ldg4(t.p1 + c); }
    if (t.mode == B200SP_VT_DY) p.p2 = ldg4(t.p2 + c);
    return p;
}
template <typename T>
__device__ __forceinline__ float4 vt_fetch(const b200sp_vtensor& t, const VtParams& p, size_t off) {
    float4 x = Vec4<T>::ld(reinterpret_cast<const T*>(t.x) + off);
    if (p.mode == B200SP_VT_PLAIN) return x;
    if (p.mode == B200SP_VT_BNACT) return f4act(f4fma(x, p.p0, p.p1), p.act);
    float4 y = Vec4<T>::ld(reinterpret_cast<const T*>(t.x2) + off);
    return f4fma(p.p0, x, f4fma(p.p1, y, p.p2));
}

// ------------------------------------------------------------------------------------------------
// Forward.  Each thread owns 4 channels and produces OW consecutive outputs per group from a
// 3 x ((OW-1)*S+3) register tile whose loads are all issued before the first FMA (memory-level
// parallelism: 15-18 independent 16-byte loads per thread in flight).
template <typename T, int S>
__global__ void __launch_bounds__(DW_NT, 2) dw_fwd_kernel(const b200sp_vtensor x, const float* __restrict__ w9c, T* __restrict__ y,
                                                          const b200sp_bnfwd bn, const int has_bn, const DwGeom gm) {
    constexpr int OW = S == 1 ? 4 : 2;
    constexpr int NC = (OW - 1) * S + 3;
    __shared__ __align__(16) float s_w[9][DW_NT];
    __shared__ float s_sum[DW_NT], s_sq[DW_NT];
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const bool active = sl < gm.SPC;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int t = 0; t < 9; ++t) s_w[t][tid] = (tid < gm.CB * 4) ? __ldg(w9c + (size_t)t * gm.C + cbase + tid) : 0.f;
    s_sum[tid] = 0.f; s_sq[tid] = 0.f;
    __syncthreads();

    float4 lsum = f4zero(), lsq = f4zero();
    if (active) {
        const VtParams xp = vt_params(x, c);
        for (long long sg = blockIdx.x; sg * gm.SPC < gm.nstrips; sg += gridDim.x) {
            const long long strip = sg * gm.SPC + sl;
            if (strip >= gm.nstrips) break;
            const int seg = (int)(strip % gm.nseg);
            const long long t2 = strip / gm.nseg;
            const int ho = (int)(t2 % gm.Ho), b = (int)(t2 / gm.Ho);
            const int wo0 = seg * gm.SEGW, wo1 = min(gm.Wo, wo0 + gm.SEGW);
            const int hi0 = ho * S - 1;
            for (int wg = wo0; wg < wo1; wg += OW) {
                float4 in[3][NC];
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const int hi = hi0 + kh, wi = wg * S - 1 + j;
                        float4 v = f4zero();
                        if (hi >= 0 && hi < gm.H && wi >= 0 && wi < gm.W)
                            v = vt_fetch<T>(x, xp, ((size_t)(b * gm.H + hi) * gm.W + wi) * gm.C + c);
                        in[kh][j] = v;
                    }
#pragma unroll
                for (int o = 0; o < OW; ++o) {
                    if (wg + o >= wo1) break;
                    float4 acc = f4zero();
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            acc = f4fma(in[kh][o * S + kw], *reinterpret_cast<const float4*>(&s_w[kh * 3 + kw][cl * 4]), acc);
                    Vec4<T>::st(y + ((size_t)(b * gm.Ho + ho) * gm.Wo + wg + o) * gm.C + c, acc);
                    lsum = f4add(lsum, acc);
                    lsq = f4fma(acc, acc, lsq);
                }
            }
        }
    }
    if (!has_bn) return;
    if (active) {
        atomicAdd(&s_sum[cl * 4 + 0], lsum.x); atomicAdd(&s_sum[cl * 4 + 1], lsum.y);
        atomicAdd(&s_sum[cl * 4 + 2], lsum.z); atomicAdd(&s_sum[cl * 4 + 3], lsum.w);
        atomicAdd(&s_sq[cl * 4 + 0], lsq.x); atomicAdd(&s_sq[cl * 4 + 1], lsq.y);
        atomicAdd(&s_sq[cl * 4 + 2], lsq.z); atomicAdd(&s_sq[cl * 4 + 3], lsq.w);
    }
    __syncthreads();
    if (tid < gm.CB * 4) {
        atomicAdd(bn.sum + cbase + tid, (double)s_sum[tid]);
        atomicAdd(bn.sumsq + cbase + tid, (double)s_sq[tid]);
    }
    if (grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_fwd_finalize_all(bn, gm.C, gm.count, tid, DW_NT);
}

// ------------------------------------------------------------------------------------------------
// Backward.  Threads walk INPUT pixels (hi, wi) in groups (2 for stride 1, 4 for stride 2); the dy
// values the group needs are fetched as one register tile up front.  dy is a virtual tensor in
// output coordinates (BatchNorm backward folded into the load).
struct DwBwdCtx {
    float4 sc, sh;
    bool has_bn, do_stats, same_src;
    ActP act;
};

template <typename T>
__device__ __forceinline__ void dw_bwd_pixel(const b200sp_vtensor& x, const VtParams& xp, const b200sp_bnbwd& bn, const DwBwdCtx& cx,
                                             const T* __restrict__ skip, T* __restrict__ g_in, size_t off, float4 dg,
                                             float4 yin, float4 a_known, float4& ls1, float4& ls2) {
    (void)x; (void)xp; (void)a_known;
    if (skip) dg = f4add(dg, Vec4<T>::ld(skip + off));
    if (cx.has_bn) {
        const float4 z = f4fma(yin, cx.sc, cx.sh);
        dg = f4mul(dg, f4actbwd(z, cx.act));
        if (cx.do_stats) { ls1 = f4add(ls1, dg); ls2 = f4fma(dg, yin, ls2); }
    }
    if (g_in) Vec4<T>::st(g_in + off, dg);
}

template <typename T, int S>
__global__ void __launch_bounds__(DW_NT, 1) dw_bwd_kernel(const b200sp_vtensor dy, const b200sp_vtensor x, const float* __restrict__ w9c,
                                                          const T* __restrict__ skip, T* __restrict__ g_in, float* __restrict__ dw9c,
                                                          const b200sp_bnbwd bn, const int has_bn, const DwGeom gm) {
    __shared__ __align__(16) float s_w[9][DW_NT];       // weights of this CTA's channels  [tap][cl*4+j]
    __shared__ float s_dw[9][DW_NT];                    // wgrad accumulators
    __shared__ float s_s1[DW_NT], s_s2[DW_NT];
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const bool active = sl < gm.SPC;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int t = 0; t < 9; ++t) {
        s_dw[t][tid] = 0.f;
        s_w[t][tid] = (tid < gm.CB * 4) ? __ldg(w9c + (size_t)t * gm.C + cbase + tid) : 0.f;
    }
    s_s1[tid] = 0.f; s_s2[tid] = 0.f;
    __syncthreads();

    DwBwdCtx cx;
    cx.has_bn = has_bn != 0;
    cx.do_stats = cx.has_bn && bn.s1 != nullptr;
    cx.act = act_params(bn.act);
    cx.sc = make_float4(1.f, 1.f, 1.f, 1.f); cx.sh = f4zero();
    if (active) {
        const VtParams dp = vt_params(dy, c);
        const VtParams xp = vt_params(x, c);
        if (cx.has_bn && bn.scale) { cx.sc = ldg4(bn.scale + c); cx.sh = ldg4(bn.shift + c); }
        cx.same_src = cx.has_bn && x.mode == B200SP_VT_BNACT && x.x == bn.y;
        float4 dwacc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) dwacc[t] = f4zero();
        float4 ls1 = f4zero(), ls2 = f4zero();       // sum g, sum g*y  (xhat applied at flush)
        const T* ybn = reinterpret_cast<const T*>(bn.y);
        auto wtap = [&](int tap) { return *reinterpret_cast<const float4*>(&s_w[tap][cl * 4]); };
        // value of the conv input (for wgrad) at element offset `off`, given the raw BN input yin
        auto in_val = [&](size_t off, float4 yin) {
            if (cx.same_src) return f4act(f4fma(yin, cx.sc, cx.sh), xp.act);
            return vt_fetch<T>(x, xp, off);
        };

        for (long long sg = blockIdx.x; sg * gm.SPC < gm.nstrips; sg += gridDim.x) {
            const long long strip = sg * gm.SPC + sl;
            if (strip >= gm.nstrips) break;
            const int seg = (int)(strip % gm.nseg);
            const long long t2 = strip / gm.nseg;
            const int hi = (int)(t2 % gm.H), b = (int)(t2 / gm.H);
            const int wi0 = seg * gm.SEGW, wi1 = min(gm.W, wi0 + gm.SEGW);
            if (S == 1) {
                constexpr int OW = 2, NC = OW + 2;
                for (int wg = wi0; wg < wi1; wg += OW) {
                    float4 D[3][NC];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int j = 0; j < NC; ++j) {
                            const int ho = hi - 1 + r, wo = wg - 1 + j;
                            float4 v = f4zero();
                            if (ho >= 0 && ho < gm.Ho && wo >= 0 && wo < gm.Wo)
                                v = vt_fetch<T>(dy, dp, ((size_t)(b * gm.Ho + ho) * gm.Wo + wo) * gm.C + c);
                            D[r][j] = v;
                        }
                    float4 yin[OW], av[OW];
                    size_t offs[OW];
#pragma unroll
                    for (int o = 0; o < OW; ++o) {
                        offs[o] = ((size_t)(b * gm.H + hi) * gm.W + min(wg + o, gm.W - 1)) * gm.C + c;
                        yin[o] = cx.has_bn ? Vec4<T>::ld(ybn + offs[o]) : f4zero();
                        av[o] = in_val(offs[o], yin[o]);
                    }
#pragma unroll
                    for (int o = 0; o < OW; ++o) {
                        if (wg + o >= wi1) break;
                        float4 dg = f4zero();
                        // output (hi+dh, wi+dw) used tap (kh,kw) = (1-dh, 1-dw);  D[r][o+s] holds dh=r-1, dw=s-1
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int sx = 0; sx < 3; ++sx) {
                                const int tap = (2 - r) * 3 + (2 - sx);
                                dg = f4fma(D[r][o + sx], wtap(tap), dg);
                                dwacc[tap] = f4fma(av[o], D[r][o + sx], dwacc[tap]);
                            }
                        dw_bwd_pixel<T>(x, xp, bn, cx, skip, g_in, offs[o], dg, yin[o], av[o], ls1, ls2);
                    }
                }
            } else {
                // stride 2: 4 input pixels starting at an even column; contributing outputs are rows
                // {hi/2} (hi even, kh=1) or {(hi+1)/2 (kh=0), (hi-1)/2 (kh=2)} (hi odd), cols wi0/2 .. wi0/2+2
                const bool odd = hi & 1;
                const int ho0 = odd ? (hi + 1) >> 1 : hi >> 1, ho1 = (hi - 1) >> 1;
                for (int wg = wi0; wg < wi1; wg += 4) {
                    float4 D[2][3];
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int ho = r == 0 ? ho0 : ho1, wo = (wg >> 1) + j;
                            float4 v = f4zero();
                            if ((r == 0 || odd) && ho < gm.Ho && wo < gm.Wo)
                                v = vt_fetch<T>(dy, dp, ((size_t)(b * gm.Ho + ho) * gm.Wo + wo) * gm.C + c);
                            D[r][j] = v;
                        }
                    float4 yin[4], av[4];
                    size_t offs[4];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        offs[o] = ((size_t)(b * gm.H + hi) * gm.W + min(wg + o, gm.W - 1)) * gm.C + c;
                        yin[o] = cx.has_bn ? Vec4<T>::ld(ybn + offs[o]) : f4zero();
                        av[o] = in_val(offs[o], yin[o]);
                    }
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (wg + o >= wi1) break;
                        float4 dg = f4zero();
                        // taps: even pixel -> (kw=1, j=o/2); odd pixel -> (kw=0, j=(o+1)/2) and (kw=2, j=(o-1)/2)
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            if (odd) {                         // r=0 -> kh=0, r=1 -> kh=2
                                const int kh = r * 2;
                                if ((o & 1) == 0) {
                                    dg = f4fma(D[r][o / 2], wtap(kh * 3 + 1), dg);
                                    if (r == 0) dwacc[1] = f4fma(av[o], D[r][o / 2], dwacc[1]); else dwacc[7] = f4fma(av[o], D[r][o / 2], dwacc[7]);
                                } else {
                                    dg = f4fma(D[r][(o + 1) / 2], wtap(kh * 3 + 0), dg);
                                    dg = f4fma(D[r][(o - 1) / 2], wtap(kh * 3 + 2), dg);
                                    if (r == 0) { dwacc[0] = f4fma(av[o], D[r][(o + 1) / 2], dwacc[0]); dwacc[2] = f4fma(av[o], D[r][(o - 1) / 2], dwacc[2]); }
                                    else        { dwacc[6] = f4fma(av[o], D[r][(o + 1) / 2], dwacc[6]); dwacc[8] = f4fma(av[o], D[r][(o - 1) / 2], dwacc[8]); }
                                }
                            } else if (r == 0) {               // kh = 1
                                if ((o & 1) == 0) {
                                    dg = f4fma(D[0][o / 2], wtap(4), dg);
                                    dwacc[4] = f4fma(av[o], D[0][o / 2], dwacc[4]);
                                } else {
                                    dg = f4fma(D[0][(o + 1) / 2], wtap(3), dg);
                                    dg = f4fma(D[0][(o - 1) / 2], wtap(5), dg);
                                    dwacc[3] = f4fma(av[o], D[0][(o + 1) / 2], dwacc[3]);
                                    dwacc[5] = f4fma(av[o], D[0][(o - 1) / 2], dwacc[5]);
                                }
                            }
                        }
                        dw_bwd_pixel<T>(x, xp, bn, cx, skip, g_in, offs[o], dg, yin[o], av[o], ls1, ls2);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            atomicAdd(&s_dw[t][cl * 4 + 0], dwacc[t].x); atomicAdd(&s_dw[t][cl * 4 + 1], dwacc[t].y);
            atomicAdd(&s_dw[t][cl * 4 + 2], dwacc[t].z); atomicAdd(&s_dw[t][cl * 4 + 3], dwacc[t].w);
        }
        if (cx.do_stats) {
            // s2 = sum g*xhat = rstd * (sum g*y - mean * sum g)
            const float4 mu = ldg4(bn.mean + c), rs = ldg4(bn.rstd + c);
            atomicAdd(&s_s1[cl * 4 + 0], ls1.x); atomicAdd(&s_s1[cl * 4 + 1], ls1.y);
            atomicAdd(&s_s1[cl * 4 + 2], ls1.z); atomicAdd(&s_s1[cl * 4 + 3], ls1.w);
            atomicAdd(&s_s2[cl * 4 + 0], rs.x * (ls2.x - mu.x * ls1.x)); atomicAdd(&s_s2[cl * 4 + 1], rs.y * (ls2.y - mu.y * ls1.y));
            atomicAdd(&s_s2[cl * 4 + 2], rs.z * (ls2.z - mu.z * ls1.z)); atomicAdd(&s_s2[cl * 4 + 3], rs.w * (ls2.w - mu.w * ls1.w));
        }
    }
    __syncthreads();
    if (tid < gm.CB * 4) {
        for (int t = 0; t < 9; ++t) atomicAdd(dw9c + (size_t)t * gm.C + cbase + tid, s_dw[t][tid]);
        if (cx.do_stats) {
            atomicAdd(bn.s1 + cbase + tid, (double)s_s1[tid]);
            atomicAdd(bn.s2 + cbase + tid, (double)s_s2[tid]);
        }
    }
    if (cx.do_stats && grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_bwd_finalize_all(bn, gm.C, gm.count, tid, DW_NT);
}

int dw_geom(DwGeom& gm, dim3& grid, int B, int H, int W, int C, int stride, bool over_input) {
    if (C % 4 || (stride != 1 && stride != 2)) return B200SP_EINVAL;
    gm.B = B; gm.H = H; gm.W = W; gm.C = C;
    gm.Ho = (H - 1) / stride + 1; gm.Wo = (W - 1) / stride + 1;
    const int C4 = C / 4;
    int cb = 1;
    for (int d = 1; d <= 64 && d <= C4; ++d) if (C4 % d == 0) cb = d;
    gm.CB = cb; gm.SPC = DW_NT / cb;
    const int rows = over_input ? H : gm.Ho, width = over_input ? W : gm.Wo;
    gm.SEGW = width <= 16 ? width : 16;
    gm.nseg = ceil_div(width, gm.SEGW);
    gm.nstrips = (long long)B * rows * gm.nseg;
    gm.count = (double)B * (over_input ? (double)H * W : (double)gm.Ho * gm.Wo);
    const int gy = C4 / cb;
    long long passes = (gm.nstrips + gm.SPC - 1) / gm.SPC;
    long long cap = (NUM_SMS * 8 + gy - 1) / gy;
    long long per = (passes + cap - 1) / cap;              // passes per CTA, balanced
    if (per < 1) per = 1;
    grid = dim3((unsigned)((passes + per - 1) / per), gy, 1);
    if (grid.x == 0) grid.x = 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Stem: x NCHW float [B,3,H,W] -> y NHWC [B,H/2,W/2,32]; one thread per output pixel, 32 channels.
constexpr int STEM_C = 32;
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                                                       const b200sp_bnfwd bn, const int has_bn, int B, int H, int W, int Ho, int Wo) {
    pdl_wait();
    pdl_trigger();
    __shared__ float s_w[27][STEM_C];            // [tap][cout]
    __shared__ float s_t[256][STEM_C + 1];       // staged outputs for the column sums
    __shared__ float s_acc[2][8][STEM_C];
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * STEM_C; i += 256) { int co = i / 27, t = i % 27; s_w[t][co] = w[i]; }
    __syncthreads();
    const long long npix = (long long)B * Ho * Wo;
    float tsum = 0.f, tsq = 0.f;                  // per (channel = tid%32, part = tid/32)
    const long long iters = (npix + 256LL * gridDim.x - 1) / (256LL * gridDim.x);
    for (long long it = 0; it < iters; ++it) {
        const long long pix = (it * gridDim.x + blockIdx.x) * 256LL + tid;
        const bool ok = pix < npix;
        float acc[STEM_C];
#pragma unroll
        for (int i = 0; i < STEM_C; ++i) acc[i] = 0.f;
        if (ok) {
            const int wo = (int)(pix % Wo);
            const long long t2 = pix / Wo;
            const int ho = (int)(t2 % Ho), b = (int)(t2 / Ho);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int hi = ho * 2 - 1 + kh, wi = wo * 2 - 1 + kw;
                        float v = 0.f;
                        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + ((size_t)(b * 3 + ci) * H + hi) * W + wi);
                        const int t = ci * 9 + kh * 3 + kw;
#pragma unroll
                        for (int co = 0; co < STEM_C; ++co) acc[co] = fmaf(v, s_w[t][co], acc[co]);
                    }
            T* yp = y + (size_t)pix * STEM_C;
#pragma unroll
            for (int i = 0; i < STEM_C; i += 4) Vec4<T>::st(yp + i, make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
        }
        if (has_bn) {
#pragma unroll
            for (int i = 0; i < STEM_C; ++i) s_t[tid][i] = ok ? acc[i] : 0.f;
            __syncthreads();
            const int ch = tid & 31, part = tid >> 5;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) { float v = s_t[part * 32 + r][ch]; tsum += v; tsq = fmaf(v, v, tsq); }
            __syncthreads();
        }
    }
    if (!has_bn) return;
    s_acc[0][tid >> 5][tid & 31] = tsum;
    s_acc[1][tid >> 5][tid & 31] = tsq;
    __syncthreads();
    if (tid < STEM_C) {
        double a = 0.0, q = 0.0;
        for (int p = 0; p < 8; ++p) { a += (double)s_acc[0][p][tid]; q += (double)s_acc[1][p][tid]; }
        atomicAdd(bn.sum + tid, a);
        atomicAdd(bn.sumsq + tid, q);
    }
    if (grid_last_cta(bn.ticket, gridDim.x))
        if (tid < STEM_C) bn_fwd_finalize_channel(bn, tid, (double)npix);
}

// dW[32][27] += sum_pix dY[pix][32] * xcol[pix][27]
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const b200sp_vtensor dy, float* __restrict__ dw,
                                                         int B, int H, int W, int Ho, int Wo) {
    __shared__ float s_dy[64][STEM_C + 1];
    __shared__ float s_x[64][28];
    const int tid = threadIdx.x;
    const int n = tid & 31, tg = tid >> 5;          // thread owns dW[n][tg], [tg+8], [tg+16], [tg+24 (<27)]
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const long long npix = (long long)B * Ho * Wo;
    for (long long base = (long long)blockIdx.x * 64; base < npix; base += (long long)gridDim.x * 64) {
        // stage 64 pixels: dy (transformed) and the 27-tap input patch
        for (int i = tid; i < 64 * 8; i += 256) {
            const int p = i >> 3, c4 = (i & 7) * 4;
            const long long pix = base + p;
            float4 v = f4zero();
            if (pix < npix) v = vt_load4<T>(dy, (size_t)pix * STEM_C + c4, c4);
            s_dy[p][c4] = v.x; s_dy[p][c4 + 1] = v.y; s_dy[p][c4 + 2] = v.z; s_dy[p][c4 + 3] = v.w;
        }
        for (int i = tid; i < 64 * 27; i += 256) {
            const int p = i / 27, t = i % 27;
            const long long pix = base + p;
            float v = 0.f;
            if (pix < npix) {
                const int wo = (int)(pix % Wo);
                const long long t2 = pix / Wo;
                const int ho = (int)(t2 % Ho), b = (int)(t2 / Ho);
                const int ci = t / 9, kh = (t % 9) / 3, kw = t % 3;
                const int hi = ho * 2 - 1 + kh, wi = wo * 2 - 1 + kw;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + ((size_t)(b * 3 + ci) * H + hi) * W + wi);
            }
            s_x[p][t] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < 64; ++p) {
            const float d = s_dy[p][n];
            acc[0] = fmaf(d, s_x[p][tg], acc[0]);
            acc[1] = fmaf(d, s_x[p][tg + 8], acc[1]);
            acc[2] = fmaf(d, s_x[p][tg + 16], acc[2]);
            if (tg + 24 < 27) acc[3] = fmaf(d, s_x[p][tg + 24], acc[3]);
        }
        __syncthreads();
    }
    atomicAdd(dw + n * 27 + tg, acc[0]);
    atomicAdd(dw + n * 27 + tg + 8, acc[1]);
    atomicAdd(dw + n * 27 + tg + 16, acc[2]);
    if (tg + 24 < 27) atomicAdd(dw + n * 27 + tg + 24, acc[3]);
}

// Second generation of the stem weight gradient (round 2; the first took 199 us = 0.14 of the copy peak because every one of
// the 27 x 64 staged input taps cost a div/mod chain and an uncoalesced scalar load).  One CTA iteration = one output row of one
// image: the 3 x 3 input rows it needs are loaded ONCE with coalesced 16-byte loads (zero rows / the left zero column stand in
// for the padding), the transformed gradient row likewise, and every thread then owns dW[n][tap], tap = tg + 8 j, j < 4, with
// the gradient read conflict-free (consecutive n) and the input broadcast (one address per warp).
constexpr int SW_XP = 232;                      // padded row pitch of the staged input rows (225 used: column wi + 1)
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad2_kernel(const float* __restrict__ x, const b200sp_vtensor dy, float* __restrict__ dw,
                                                          int B, int H, int W, int Ho, int Wo) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float sw_smem[];
    float* s_x = sw_smem;                        // [3 ci][3 kh][SW_XP]
    float* s_dy = sw_smem + 9 * SW_XP;           // [Wo][32]
    const int tid = threadIdx.x;
    // register tile: 4 output channels x 4 taps per thread (5 shared-memory loads per 16 FMAs -- the first cut of this kernel,
    // 1 channel x 4 taps, was bound by its 5 loads per 4 FMAs); 4 pixel groups of 64 threads share a row
    const int grp = tid >> 6, l = tid & 63, nb = l & 7, tb = l >> 3;
    int xo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = min(tb * 4 + j, 26);
        xo[j] = ((t / 9) * 3 + (t % 9) / 3) * SW_XP + (t % 3);      // + 2*wo: column (2 wo - 1 + kw) + 1
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int i = tid; i < 9; i += 256) s_x[i * SW_XP] = 0.f;       // the zero column left of the image (wi = -1)
    const int W4 = W >> 2, nrow = B * Ho;
    for (int r = blockIdx.x; r < nrow; r += gridDim.x) {
        const int b = r / Ho, ho = r - b * Ho;
        __syncthreads();                                             // previous row fully consumed
        for (int i = tid; i < 9 * W4; i += 256) {
            const int row = i / W4, c4 = (i - row * W4) * 4;
            const int ci = row / 3, hi = 2 * ho - 1 + (row - ci * 3);
            float4 v = f4zero();
            if (hi >= 0 && hi < H) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)(b * 3 + ci) * H + hi) * W + c4));
            float* d = s_x + row * SW_XP + 1 + c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        const size_t prow = (size_t)r * Wo * STEM_C;
        for (int i = tid; i < Wo * (STEM_C / 4); i += 256) {
            const int c4 = (i & (STEM_C / 4 - 1)) * 4;
            const float4 v = vt_load4<T>(dy, prow + (size_t)i * 4, c4);
            *reinterpret_cast<float4*>(s_dy + i * 4) = v;
        }
        __syncthreads();
        if (tb < 7) {
#pragma unroll 2
            for (int p = grp; p < Wo; p += 4) {
                const float4 d = *reinterpret_cast<const float4*>(s_dy + p * STEM_C + nb * 4);
                const float* xr = s_x + 2 * p;
                const float x0 = xr[xo[0]], x1 = xr[xo[1]], x2 = xr[xo[2]], x3 = xr[xo[3]];
                acc[0][0] = fmaf(d.x, x0, acc[0][0]); acc[0][1] = fmaf(d.x, x1, acc[0][1]); acc[0][2] = fmaf(d.x, x2, acc[0][2]); acc[0][3] = fmaf(d.x, x3, acc[0][3]);
                acc[1][0] = fmaf(d.y, x0, acc[1][0]); acc[1][1] = fmaf(d.y, x1, acc[1][1]); acc[1][2] = fmaf(d.y, x2, acc[1][2]); acc[1][3] = fmaf(d.y, x3, acc[1][3]);
                acc[2][0] = fmaf(d.z, x0, acc[2][0]); acc[2][1] = fmaf(d.z, x1, acc[2][1]); acc[2][2] = fmaf(d.z, x2, acc[2][2]); acc[2][3] = fmaf(d.z, x3, acc[2][3]);
                acc[3][0] = fmaf(d.w, x0, acc[3][0]); acc[3][1] = fmaf(d.w, x1, acc[3][1]); acc[3][2] = fmaf(d.w, x2, acc[3][2]); acc[3][3] = fmaf(d.w, x3, acc[3][3]);
            }
        }
    }
    // reduce the four pixel groups in shared memory, then one atomic per weight and CTA
    __syncthreads();
    float* s_red = s_dy;                         // [4 groups][32 x 28]   (Wo*32 >= 4*896 floats is checked by the launcher)
    if (tb < 7) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s_red[grp * 896 + (nb * 4 + i) * 28 + tb * 4 + j] = acc[i][j];
    }
    __syncthreads();
    for (int o = tid; o < 32 * 27; o += 256) {
        const int n = o / 27, t = o - n * 27;
        const int q = n * 28 + t;
        atomicAdd(dw + o, s_red[q] + s_red[896 + q] + s_red[2 * 896 + q] + s_red[3 * 896 + q]);
    }
}

}  // namespace

int dw_fwd_legacy(const b200sp_vtensor* x, const float* w9c, void* y, const b200sp_bnfwd* bn,
                             int B, int H, int W, int C, int stride, int dtype, void* stream) {
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    if (!x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    DwGeom gm; dim3 grid;
    if (int rc = dw_geom(gm, grid, B, H, W, C, stride, false)) return rc;
    b200sp_bnfwd b = {};
    if (bn) b = *bn;
    if (stride == 1) dw_fwd_kernel<float, 1><<<grid, DW_NT, 0, (cudaStream_t)stream>>>(*x, w9c, (float*)y, b, bn != nullptr, gm);
    else             dw_fwd_kernel<float, 2><<<grid, DW_NT, 0, (cudaStream_t)stream>>>(*x, w9c, (float*)y, b, bn != nullptr, gm);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

int dw_bwd_legacy(const b200sp_vtensor* dy, const b200sp_vtensor* x, const float* w9c, const void* skip,
                             void* g_in, float* dw9c, const b200sp_bnbwd* bn,
                             int B, int H, int W, int C, int stride, int dtype, void* stream) {
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    if (!dy || !x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    DwGeom gm; dim3 grid;
    if (int rc = dw_geom(gm, grid, B, H, W, C, stride, true)) return rc;
    b200sp_vtensor d = *dy;
    if (d.mode != B200SP_VT_DY) d.x2 = d.x;
    b200sp_bnbwd b = {};
    if (bn) b = *bn;
    if (stride == 1) dw_bwd_kernel<float, 1><<<grid, DW_NT, 0, (cudaStream_t)stream>>>(d, *x, w9c, (const float*)skip, (float*)g_in, dw9c, b, bn != nullptr, gm);
    else             dw_bwd_kernel<float, 2><<<grid, DW_NT, 0, (cudaStream_t)stream>>>(d, *x, w9c, (const float*)skip, (float*)g_in, dw9c, b, bn != nullptr, gm);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_stem_fwd(const float* x_nchw, const float* w, void* y, const b200sp_bnfwd* bn,
                               int B, int H, int W, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    int grid = (int)((npix + 255) / 256);
    if (grid > NUM_SMS * 4) grid = NUM_SMS * 4;
    b200sp_bnfwd b = {};
    if (bn) b = *bn;
    if (dtype == B200SP_F32) b200sp_launch_pdl(stem_fwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x_nchw, w, (float*)y, b, (int)(bn != nullptr), B, H, W, Ho, Wo);
    else stem_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, w, (bf16*)y, b, bn != nullptr, B, H, W, Ho, Wo);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_stem_wgrad(const float* x_nchw, const b200sp_vtensor* dy, float* dw,
                                 int B, int H, int W, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    int grid = (int)((npix + 63) / 64);
    if (grid > NUM_SMS * 8) grid = NUM_SMS * 8;
    static int v2 = -1;
    if (v2 < 0) { const char* e = getenv("B200SP_STEM_WGRAD"); v2 = (e && e[0] == '1') ? 0 : 1; }
    if (v2 && dtype == B200SP_F32 && W % 4 == 0 && W + 1 <= SW_XP && Wo * STEM_C >= 4 * 896 && ((uintptr_t)x_nchw & 15) == 0) {
        const size_t smem = sizeof(float) * (9 * SW_XP + (size_t)Wo * STEM_C);
        int g2 = B * Ho < NUM_SMS * 4 ? B * Ho : NUM_SMS * 4;
        b200sp_launch_pdl(stem_wgrad2_kernel<float>, dim3(g2), dim3(256), smem, (cudaStream_t)stream, x_nchw, *dy, dw, B, H, W, Ho, Wo);
        B200SP_COUNT_LAUNCH();
        B200SP_RETURN_LAST();
    }
    if (dtype == B200SP_F32) stem_wgrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, *dy, dw, B, H, W, Ho, Wo);
    else stem_wgrad_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, *dy, dw, B, H, W, Ho, Wo);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
