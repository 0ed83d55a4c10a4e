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
__device__ __forceinline__ float4 f4act(float4 z, int act) {
    return make_float4(act_fwd(z.x, act), act_fwd(z.y, act), act_fwd(z.z, act), act_fwd(z.w, act));
}
__device__ __forceinline__ float4 f4actbwd(float4 z, int act) {
    return make_float4(act_bwd(z.x, act), act_bwd(z.y, act), act_bwd(z.z, act), act_bwd(z.w, act));
}

// per-thread copy of a virtual tensor's channel parameters (4 channels)
struct VtParams {
    float4 p0, p1, p2;
    int mode, act;
};
__device__ __forceinline__ VtParams vt_params(const b200sp_vtensor& t, int c) {
    VtParams p;
    p.mode = t.mode; p.act = t.act;
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
template <typename T, int S>
__global__ void __launch_bounds__(DW_NT) dw_fwd_kernel(const b200sp_vtensor x, const float* __restrict__ w9c, T* __restrict__ y,
                                                       const b200sp_bnfwd bn, const int has_bn, const DwGeom gm) {
    __shared__ float s_sum[DW_NT], s_sq[DW_NT];
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const bool active = sl < gm.SPC;
    const int c = (blockIdx.y * gm.CB + cl) * 4;
    s_sum[tid] = 0.f; s_sq[tid] = 0.f;
    __syncthreads();

    float4 lsum = f4zero(), lsq = f4zero();
    if (active) {
        float4 w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = ldg4(w9c + (size_t)t * gm.C + c);
        const VtParams xp = vt_params(x, c);
        for (long long sg = blockIdx.x; sg * gm.SPC < gm.nstrips; sg += gridDim.x) {
            const long long strip = sg * gm.SPC + sl;
            if (strip >= gm.nstrips) break;
            const int seg = (int)(strip % gm.nseg);
            const long long t2 = strip / gm.nseg;
            const int ho = (int)(t2 % gm.Ho), b = (int)(t2 / gm.Ho);
            const int wo0 = seg * gm.SEGW, wo1 = min(gm.Wo, wo0 + gm.SEGW);
            const int hi0 = ho * S - 1;
            float4 win[3][3];
            auto load_col = [&](int wi, int slot) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hi = hi0 + kh;
                    float4 v = f4zero();
                    if (hi >= 0 && hi < gm.H && wi >= 0 && wi < gm.W)
                        v = vt_fetch<T>(x, xp, ((size_t)(b * gm.H + hi) * gm.W + wi) * gm.C + c);
                    win[kh][slot] = v;
                }
            };
            if (S == 1) { load_col(wo0 - 1, 1); load_col(wo0, 2); }
            else        { load_col(wo0 * 2 - 1, 2); }
            for (int wo = wo0; wo < wo1; ++wo) {
                if (S == 1) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) { win[kh][0] = win[kh][1]; win[kh][1] = win[kh][2]; }
                    load_col(wo + 1, 2);
                } else {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) win[kh][0] = win[kh][2];
                    load_col(wo * 2, 1);
                    load_col(wo * 2 + 1, 2);
                }
                float4 acc = f4zero();
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) acc = f4fma(win[kh][kw], w[kh * 3 + kw], acc);
                Vec4<T>::st(y + ((size_t)(b * gm.Ho + ho) * gm.Wo + wo) * gm.C + c, acc);
                lsum = f4add(lsum, acc);
                lsq = f4fma(acc, acc, lsq);
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
        const int cc = blockIdx.y * gm.CB * 4 + tid;
        atomicAdd(bn.sum + cc, (double)s_sum[tid]);
        atomicAdd(bn.sumsq + cc, (double)s_sq[tid]);
    }
    if (grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        for (int cc = tid; cc < gm.C; cc += DW_NT) bn_fwd_finalize_channel(bn, cc, gm.count);
}

// ------------------------------------------------------------------------------------------------
// Backward.  Threads walk INPUT pixels (hi, wi).  dy is a virtual tensor in output coordinates.
template <typename T, int S>
__global__ void __launch_bounds__(DW_NT) dw_bwd_kernel(const b200sp_vtensor dy, const b200sp_vtensor x, const float* __restrict__ w9c,
                                                       const T* __restrict__ skip, T* __restrict__ g_in, float* __restrict__ dw9c,
                                                       const b200sp_bnbwd bn, const int has_bn, const DwGeom gm) {
    __shared__ __align__(16) float s_w[9][DW_NT];       // weights of this CTA's channels  [tap][cl*4+j]
    __shared__ float s_dw[9][DW_NT];      // wgrad accumulators
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

    const bool do_stats = has_bn && bn.s1 != nullptr;
    if (active) {
        const VtParams dp = vt_params(dy, c);
        const VtParams xp = vt_params(x, c);
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = f4zero(), mu = f4zero(), rs = f4zero();
        if (has_bn && bn.scale) { sc = ldg4(bn.scale + c); sh = ldg4(bn.shift + c); }
        if (do_stats) { mu = ldg4(bn.mean + c); rs = ldg4(bn.rstd + c); }
        const bool same_src = has_bn && x.mode == B200SP_VT_BNACT && x.x == bn.y;
        float4 dwacc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) dwacc[t] = f4zero();
        float4 ls1 = f4zero(), ls2 = f4zero();

        for (long long sg = blockIdx.x; sg * gm.SPC < gm.nstrips; sg += gridDim.x) {
            const long long strip = sg * gm.SPC + sl;
            if (strip >= gm.nstrips) break;
            const int seg = (int)(strip % gm.nseg);
            const long long t2 = strip / gm.nseg;
            const int hi = (int)(t2 % gm.H), b = (int)(t2 / gm.H);
            const int wi0 = seg * gm.SEGW, wi1 = min(gm.W, wi0 + gm.SEGW);

            float4 win[3][3];   // S==1 only: dy at rows hi-1..hi+1, cols wi-1..wi+1
            auto load_col = [&](int wo, int slot) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int ho = hi - 1 + r;
                    float4 v = f4zero();
                    if (ho >= 0 && ho < gm.Ho && wo >= 0 && wo < gm.Wo)
                        v = vt_fetch<T>(dy, dp, ((size_t)(b * gm.Ho + ho) * gm.Wo + wo) * gm.C + c);
                    win[r][slot] = v;
                }
            };
            if (S == 1) { load_col(wi0 - 1, 1); load_col(wi0, 2); }

            for (int wi = wi0; wi < wi1; ++wi) {
                const size_t off = ((size_t)(b * gm.H + hi) * gm.W + wi) * gm.C + c;
                // ---- value of the conv input at this pixel (for wgrad) and its pre-activation z
                float4 yin = f4zero(), z = f4zero(), a;
                if (has_bn) { yin = Vec4<T>::ld(reinterpret_cast<const T*>(bn.y) + off); z = f4fma(yin, sc, sh); }
                if (same_src) a = f4act(z, xp.act);
                else          a = vt_fetch<T>(x, xp, off);

                float4 dg = f4zero();
                if (S == 1) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) { win[r][0] = win[r][1]; win[r][1] = win[r][2]; }
                    load_col(wi + 1, 2);
                    // output (hi+dh, wi+dw) used tap (kh,kw) = (1-dh, 1-dw);  win[r][s] holds dh=r-1, dw=s-1
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            const int tap = (2 - r) * 3 + (2 - s);
                            const float4 wv = *reinterpret_cast<const float4*>(&s_w[tap][cl * 4]);
                            dg = f4fma(win[r][s], wv, dg);
                            dwacc[tap] = f4fma(a, win[r][s], dwacc[tap]);
                        }
                } else {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int th = hi + 1 - kh;
                        if (th < 0 || (th & 1)) continue;
                        const int ho = th >> 1;
                        if (ho >= gm.Ho) continue;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const int tw = wi + 1 - kw;
                            if (tw < 0 || (tw & 1)) continue;
                            const int wo = tw >> 1;
                            if (wo >= gm.Wo) continue;
                            const float4 d = vt_fetch<T>(dy, dp, ((size_t)(b * gm.Ho + ho) * gm.Wo + wo) * gm.C + c);
                            const float4 wv = *reinterpret_cast<const float4*>(&s_w[kh * 3 + kw][cl * 4]);
                            dg = f4fma(d, wv, dg);
                            dwacc[kh * 3 + kw] = f4fma(a, d, dwacc[kh * 3 + kw]);
                        }
                    }
                }
                if (skip) dg = f4add(dg, Vec4<T>::ld(skip + off));
                if (has_bn) {
                    dg = f4mul(dg, f4actbwd(z, bn.act));
                    if (do_stats) {
                        ls1 = f4add(ls1, dg);
                        ls2 = make_float4(fmaf(dg.x, (yin.x - mu.x) * rs.x, ls2.x), fmaf(dg.y, (yin.y - mu.y) * rs.y, ls2.y),
                                          fmaf(dg.z, (yin.z - mu.z) * rs.z, ls2.z), fmaf(dg.w, (yin.w - mu.w) * rs.w, ls2.w));
                    }
                }
                if (g_in) Vec4<T>::st(g_in + off, dg);
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            atomicAdd(&s_dw[t][cl * 4 + 0], dwacc[t].x); atomicAdd(&s_dw[t][cl * 4 + 1], dwacc[t].y);
            atomicAdd(&s_dw[t][cl * 4 + 2], dwacc[t].z); atomicAdd(&s_dw[t][cl * 4 + 3], dwacc[t].w);
        }
        if (do_stats) {
            atomicAdd(&s_s1[cl * 4 + 0], ls1.x); atomicAdd(&s_s1[cl * 4 + 1], ls1.y);
            atomicAdd(&s_s1[cl * 4 + 2], ls1.z); atomicAdd(&s_s1[cl * 4 + 3], ls1.w);
            atomicAdd(&s_s2[cl * 4 + 0], ls2.x); atomicAdd(&s_s2[cl * 4 + 1], ls2.y);
            atomicAdd(&s_s2[cl * 4 + 2], ls2.z); atomicAdd(&s_s2[cl * 4 + 3], ls2.w);
        }
    }
    __syncthreads();
    if (tid < gm.CB * 4) {
        for (int t = 0; t < 9; ++t) atomicAdd(dw9c + (size_t)t * gm.C + cbase + tid, s_dw[t][tid]);
        if (do_stats) {
            atomicAdd(bn.s1 + cbase + tid, (double)s_s1[tid]);
            atomicAdd(bn.s2 + cbase + tid, (double)s_s2[tid]);
        }
    }
    if (do_stats && grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        for (int cc = tid; cc < gm.C; cc += DW_NT) bn_bwd_finalize_channel(bn, cc, gm.count);
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
    grid = dim3((unsigned)(passes < cap ? passes : cap), gy, 1);
    if (grid.x == 0) grid.x = 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Stem: x NCHW float [B,3,H,W] -> y NHWC [B,H/2,W/2,32]; one thread per output pixel, 32 channels.
constexpr int STEM_C = 32;
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                                                       const b200sp_bnfwd bn, const int has_bn, int B, int H, int W, int Ho, int Wo) {
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

}  // namespace

extern "C" int b200sp_dw_fwd(const b200sp_vtensor* x, const float* w9c, void* y, const b200sp_bnfwd* bn,
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

extern "C" int b200sp_dw_bwd(const b200sp_vtensor* dy, const b200sp_vtensor* x, const float* w9c, const void* skip,
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
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    int grid = (int)((npix + 255) / 256);
    if (grid > NUM_SMS * 4) grid = NUM_SMS * 4;
    b200sp_bnfwd b = {};
    if (bn) b = *bn;
    stem_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, w, (float*)y, b, bn != nullptr, B, H, W, Ho, Wo);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_stem_wgrad(const float* x_nchw, const b200sp_vtensor* dy, float* dw,
                                 int B, int H, int W, int dtype, void* stream) {
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long npix = (long long)B * Ho * Wo;
    int grid = (int)((npix + 63) / 64);
    if (grid > NUM_SMS * 8) grid = NUM_SMS * 8;
    stem_wgrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, *dy, dw, B, H, W, Ho, Wo);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
