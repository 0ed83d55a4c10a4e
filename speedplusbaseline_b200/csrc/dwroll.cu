// dwroll.cu -- depthwise 3x3 convolution, "rolling window" formulation (forward; fused
// dgrad + wgrad + activation/BatchNorm backward).  Replaces torch conv2d(groups=C) at
// park2019.py:47 and torchvision mobilenetv2.py:45-49 and their autograd backward.
//
// HBM-bound CUDA-core stencils over NHWC.  A thread owns 4 channels (one 16-byte vector) of a
// narrow column strip and walks DOWN the image keeping the rows the 3x3 window needs in registers,
// so every input element is fetched (and BN+activation-transformed) ~1.25x instead of 4.5x, all
// loads of a step are independent (6-12 x 16 B in flight per thread) and no shared-memory staging
// or barrier sits on the critical path.  BatchNorm statistics (forward: sum, sumsq; backward:
// sum g, sum g*xhat) are reduced in registers -> shared atomics -> one double atomic per channel per
// CTA; the last CTA finalises them (common.cuh).
#include <cstdlib>
#include "common.cuh"

int dw_fwd_legacy(const b200sp_vtensor* x, const float* w9c, void* y, const b200sp_bnfwd* bn,
                  int B, int H, int W, int C, int stride, int dtype, void* stream);
int dw_bwd_legacy(const b200sp_vtensor* dy, const b200sp_vtensor* x, const float* w9c, const void* skip,
                  void* g_in, float* dw9c, const b200sp_bnbwd* bn,
                  int B, int H, int W, int C, int stride, int dtype, void* stream);

namespace {

constexpr int RNT = 128;          // threads per CTA
#ifndef DWR_S1_OCC
#define DWR_S1_OCC 3                // resident CTAs per SM requested for the stride-1 second-generation kernels (register cap 168)
#endif
constexpr int MAXCB = 32;         // channel quads per CTA row

struct RGeom {
    int B, H, W, C, Ho, Wo;
    int CB, SPC;                  // channel quads per CTA, work items per CTA
    int nWG, nSeg, SEG;           // column groups, row segments, rows per segment
    long long nitems;
    double count;
};

__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4act(float4 z, ActP a) {
    return make_float4(act_fwd(z.x, a), act_fwd(z.y, a), act_fwd(z.z, a), act_fwd(z.w, a));
}
__device__ __forceinline__ float4 f4actbwd(float4 z, ActP a) {
    return make_float4(act_bwd(z.x, a), act_bwd(z.y, a), act_bwd(z.z, a), act_bwd(z.w, a));
}

// virtual-tensor access with the mode fixed at compile time and the per-channel parameters in
// registers.  Loads are issued UNCONDITIONALLY from clamped (always valid) addresses so that all the
// loads of a step are in flight together; padding is applied afterwards with a select.
enum { XM_PLAIN = 0, XM_BNACT = 1, XM_DY = 2 };
struct VtP { float4 a, b, c; ActP act; };
template <int MODE>
__device__ __forceinline__ VtP vtp_load(const b200sp_vtensor& t, int c) {
    VtP p;
    p.act = act_params(t.act);
    p.a = p.b = p.c = f4zero();
    if (MODE != XM_PLAIN) { p.a = ldg4(t.p0 + c); p.b = ldg4(t.p1 + c); }
    if (MODE == XM_DY) p.c = ldg4(t.p2 + c);
    return p;
}
template <int MODE>
__device__ __forceinline__ float4 vtp_apply(const VtP& p, float4 x, float4 x2, bool valid) {
    float4 v = x;
    if (MODE == XM_BNACT) v = f4act(f4fma(x, p.a, p.b), p.act);
    if (MODE == XM_DY) v = f4fma(p.a, x, f4fma(p.b, x2, p.c));
    return valid ? v : f4zero();
}

struct Item { int b, r_a, r_b, wg; bool ok; };
__device__ __forceinline__ Item get_item(const RGeom& gm, int sl, int rows) {
    Item it;
    const long long item = (long long)blockIdx.x * gm.SPC + sl;
    it.ok = sl < gm.SPC && item < gm.nitems;
    const long long ii = it.ok ? item : 0;
    it.wg = (int)(ii % gm.nWG);
    const long long t = ii / gm.nWG;
    const int seg = (int)(t % gm.nSeg);
    it.b = (int)(t / gm.nSeg);
    it.r_a = seg * gm.SEG;
    it.r_b = min(rows, it.r_a + gm.SEG);
    return it;
}

// raw (untransformed) loads of one row segment; apply_row() turns them into values later, so a caller can
// put ALL the loads of a step in flight before the first use
template <typename T, int MODE, int NC>
__device__ __forceinline__ void load_row_raw(const b200sp_vtensor& t, size_t img, int row, int rows, int col0, int cols,
                                             int C, int c, float4 (&raw)[NC], float4 (&raw2)[NC]) {
    const size_t rowoff = (img + (size_t)min(max(row, 0), rows - 1) * cols) * C + c;
    const T* x = reinterpret_cast<const T*>(t.x);
    const T* x2 = reinterpret_cast<const T*>(t.x2);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const size_t off = rowoff + (size_t)min(max(col0 + j, 0), cols - 1) * C;
        raw[j] = Vec4<T>::ld(x + off);
        if (MODE == XM_DY) raw2[j] = Vec4<T>::ld(x2 + off); else raw2[j] = f4zero();
    }
}
template <int MODE, int NC>
__device__ __forceinline__ void apply_row(const VtP& p, int row, int rows, int col0, int cols, const float4 (&raw)[NC],
                                          const float4 (&raw2)[NC], float4 (&r)[NC]) {
    const bool rok = row >= 0 && row < rows;
#pragma unroll
    for (int j = 0; j < NC; ++j) r[j] = vtp_apply<MODE>(p, raw[j], raw2[j], rok && col0 + j >= 0 && col0 + j < cols);
}

// load + transform one row segment of NC pixels (4 channels each) of a virtual [rows, cols, C] image
template <typename T, int MODE, int NC>
__device__ __forceinline__ void load_row(const b200sp_vtensor& t, const VtP& p, size_t img, int row, int rows, int col0, int cols,
                                         int C, int c, float4 (&r)[NC]) {
    const bool rok = row >= 0 && row < rows;
    const size_t rowoff = (img + (size_t)min(max(row, 0), rows - 1) * cols) * C + c;
    const T* x = reinterpret_cast<const T*>(t.x);
    const T* x2 = reinterpret_cast<const T*>(t.x2);
    float4 raw[NC], raw2[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const size_t off = rowoff + (size_t)min(max(col0 + j, 0), cols - 1) * C;
        raw[j] = Vec4<T>::ld(x + off);
        if (MODE == XM_DY) raw2[j] = Vec4<T>::ld(x2 + off); else raw2[j] = f4zero();
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) r[j] = vtp_apply<MODE>(p, raw[j], raw2[j], rok && col0 + j >= 0 && col0 + j < cols);
}

// ------------------------------------------------------------------------------------------------
// Forward: thread = 4 channels x OW output columns, rolling over output rows.
template <typename T, int S, int XM>
__global__ void __launch_bounds__(RNT, 3) dwr_fwd_kernel(const b200sp_vtensor x, const float* __restrict__ w9c, T* __restrict__ y,
                                                         const b200sp_bnfwd bn, const int has_bn, const RGeom gm) {
    constexpr int OW = S == 1 ? 4 : 2;
    constexpr int NC = (OW - 1) * S + 3;
    __shared__ float4 s_w[9][MAXCB];
    __shared__ float s_sum[MAXCB * 4], s_sq[MAXCB * 4];
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int i = tid; i < 9 * gm.CB; i += RNT) s_w[i / gm.CB][i % gm.CB] = ldg4(w9c + (size_t)(i / gm.CB) * gm.C + cbase + (i % gm.CB) * 4);
    if (tid < MAXCB * 4) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();

    const Item it = get_item(gm, sl, gm.Ho);
    float4 lsum = f4zero(), lsq = f4zero();
    if (it.ok) {
        const VtP xp = vtp_load<XM>(x, c);
        const int wo0 = it.wg * OW;
        const int wi0 = wo0 * S - 1;
        const size_t img = (size_t)it.b * gm.H * gm.W;
        float4 r0[NC], r1[NC], r2[NC];
        if (S == 1) {
            load_row<T, XM, NC>(x, xp, img, it.r_a - 1, gm.H, wi0, gm.W, gm.C, c, r0);
            load_row<T, XM, NC>(x, xp, img, it.r_a, gm.H, wi0, gm.W, gm.C, c, r1);
        } else {
            load_row<T, XM, NC>(x, xp, img, 2 * it.r_a - 1, gm.H, wi0, gm.W, gm.C, c, r0);
        }
        for (int ho = it.r_a; ho < it.r_b; ++ho) {
            if (S == 1) {
                load_row<T, XM, NC>(x, xp, img, ho + 1, gm.H, wi0, gm.W, gm.C, c, r2);
            } else {
                load_row<T, XM, NC>(x, xp, img, 2 * ho, gm.H, wi0, gm.W, gm.C, c, r1);
                load_row<T, XM, NC>(x, xp, img, 2 * ho + 1, gm.H, wi0, gm.W, gm.C, c, r2);
            }
            T* yrow = y + (((size_t)it.b * gm.Ho + ho) * gm.Wo) * gm.C + c;
#pragma unroll
            for (int o = 0; o < OW; ++o) {
                float4 acc = f4mul(r0[o * S], s_w[0][cl]);
                acc = f4fma(r0[o * S + 1], s_w[1][cl], acc); acc = f4fma(r0[o * S + 2], s_w[2][cl], acc);
                acc = f4fma(r1[o * S], s_w[3][cl], acc); acc = f4fma(r1[o * S + 1], s_w[4][cl], acc); acc = f4fma(r1[o * S + 2], s_w[5][cl], acc);
                acc = f4fma(r2[o * S], s_w[6][cl], acc); acc = f4fma(r2[o * S + 1], s_w[7][cl], acc); acc = f4fma(r2[o * S + 2], s_w[8][cl], acc);
                if (wo0 + o < gm.Wo) {
                    Vec4<T>::st(yrow + (size_t)(wo0 + o) * gm.C, acc);
                    lsum = f4add(lsum, acc);
                    lsq = f4fma(acc, acc, lsq);
                }
            }
            if (S == 1) {
#pragma unroll
                for (int j = 0; j < NC; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < NC; ++j) r0[j] = r2[j];
            }
        }
    }
    if (!has_bn) return;
    if (it.ok) {
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
        bn_fwd_finalize_all(bn, gm.C, gm.count, tid, RNT);
}

// ------------------------------------------------------------------------------------------------
// Backward.  Threads own INPUT pixels and roll down the input rows; dy is a virtual tensor in output
// coordinates (BatchNorm backward folded into the load).
//   g_in = (dgrad(dy) [+ skip]) * act'(z_in);   dw9c += wgrad;   s1/s2 reductions for the input's BN.
struct BwdShared {
    float4 w[9][MAXCB];
    float dw[9][MAXCB * 4];
    float s1[MAXCB * 4], s2[MAXCB * 4];
};

struct BwdCtx {
    float4 sc, sh;
    bool has_bn, do_stats, same_src, has_skip;
    ActP act, xact;
};

// per-pixel tail: skip add, activation-derivative mask, BN-backward partial sums, store
template <typename T>
__device__ __forceinline__ void bwd_pixel_finish(const BwdCtx& cx, T* __restrict__ g_in, size_t off,
                                                 float4 dg, float4 yin, float4 skipv, float4& ls1, float4& ls2) {
    dg = f4add(dg, skipv);
    if (cx.has_bn) {
        dg = f4mul(dg, f4actbwd(f4fma(yin, cx.sc, cx.sh), cx.act));
        ls1 = f4add(ls1, dg); ls2 = f4fma(dg, yin, ls2);
    }
    if (g_in) Vec4<T>::st(g_in + off, dg);
}

template <typename T, int S, int DM, int XM>
__global__ void __launch_bounds__(RNT, 3) dwr_bwd_kernel(const b200sp_vtensor dy, const b200sp_vtensor x, const float* __restrict__ w9c,
                                                         const T* __restrict__ skip, T* __restrict__ g_in, float* __restrict__ dw9c,
                                                         const b200sp_bnbwd bn, const int has_bn, const RGeom gm) {
    __shared__ BwdShared sh;
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int i = tid; i < 9 * gm.CB; i += RNT) sh.w[i / gm.CB][i % gm.CB] = ldg4(w9c + (size_t)(i / gm.CB) * gm.C + cbase + (i % gm.CB) * 4);
    for (int i = tid; i < 9 * MAXCB * 4; i += RNT) (&sh.dw[0][0])[i] = 0.f;
    if (tid < MAXCB * 4) { sh.s1[tid] = 0.f; sh.s2[tid] = 0.f; }
    __syncthreads();

    // S == 1: rows = input rows, item covers IW = 2 input columns.
    // S == 2: rows = 2x2 input block rows (a), item covers 1 block column = 2 input columns.
    const Item it = get_item(gm, sl, S == 1 ? gm.H : gm.Ho);
    BwdCtx cx;
    cx.has_bn = has_bn != 0;
    cx.do_stats = cx.has_bn && bn.s1 != nullptr;
    cx.has_skip = skip != nullptr;
    cx.act = act_params(bn.act);
    cx.xact = act_params(x.act);
    cx.sc = make_float4(1.f, 1.f, 1.f, 1.f); cx.sh = f4zero();
    cx.same_src = cx.has_bn && XM == XM_BNACT && x.x == bn.y;
    float4 ls1 = f4zero(), ls2 = f4zero();
    if (it.ok) {
        const VtP dp = vtp_load<DM>(dy, c);
        const VtP xp = vtp_load<XM>(x, c);
        if (cx.has_bn && bn.scale) { cx.sc = ldg4(bn.scale + c); cx.sh = ldg4(bn.shift + c); }
        const T* ybn = reinterpret_cast<const T*>(bn.y);
        const T* xraw = reinterpret_cast<const T*>(x.x);
        float4 dwacc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) dwacc[t] = f4zero();
        const size_t oimg = (size_t)it.b * gm.Ho * gm.Wo, iimg = (size_t)it.b * gm.H * gm.W;
        // per input pixel: raw BN input (yin), conv-input value (av, for wgrad), skip gradient; loads are unconditional
        // raw loads first (yin, conv-input, skip gradient), the conv-input VALUE av (for wgrad) afterwards
        auto pixel_loads = [&](size_t off, float4& yin, float4& xr, float4& skv) {
            yin = f4zero(); skv = f4zero(); xr = f4zero();
            if (cx.has_bn) yin = Vec4<T>::ld(ybn + off);
            if (!cx.same_src) xr = Vec4<T>::ld(xraw + off);
            if (cx.has_skip) skv = Vec4<T>::ld(skip + off);
        };
        auto in_val = [&](float4 yin, float4 xr) {
            return cx.same_src ? f4act(f4fma(yin, cx.sc, cx.sh), cx.xact) : vtp_apply<XM>(xp, xr, f4zero(), true);
        };
        if (S == 1) {
            constexpr int IW = 2, NC = IW + 2;
            const int wi0 = it.wg * IW;
            float4 D0[NC], D1[NC], D2[NC];
            load_row<T, DM, NC>(dy, dp, oimg, it.r_a - 1, gm.Ho, wi0 - 1, gm.Wo, gm.C, c, D0);
            load_row<T, DM, NC>(dy, dp, oimg, it.r_a, gm.Ho, wi0 - 1, gm.Wo, gm.C, c, D1);
            for (int hi = it.r_a; hi < it.r_b; ++hi) {
                float4 yin[IW], av[IW], skv[IW], rg[NC], ry[NC];
                size_t offs[IW];
                load_row_raw<T, DM, NC>(dy, oimg, hi + 1, gm.Ho, wi0 - 1, gm.Wo, gm.C, c, rg, ry);
#pragma unroll
                for (int o = 0; o < IW; ++o) {
                    offs[o] = (iimg + (size_t)hi * gm.W + min(wi0 + o, gm.W - 1)) * gm.C + c;
                    pixel_loads(offs[o], yin[o], av[o], skv[o]);
                }
                apply_row<DM, NC>(dp, hi + 1, gm.Ho, wi0 - 1, gm.Wo, rg, ry, D2);
#pragma unroll
                for (int o = 0; o < IW; ++o) av[o] = in_val(yin[o], av[o]);
#pragma unroll
                for (int o = 0; o < IW; ++o) {
                    if (wi0 + o < gm.W) {
                        float4 dg = f4zero();
                        // output (hi+dh, wi+dw) used tap (kh,kw) = (1-dh, 1-dw);  D<r>[o+sx] holds dh = r-1, dw = sx-1
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            dg = f4fma(D0[o + sx], sh.w[6 + (2 - sx)][cl], dg);
                            dg = f4fma(D1[o + sx], sh.w[3 + (2 - sx)][cl], dg);
                            dg = f4fma(D2[o + sx], sh.w[0 + (2 - sx)][cl], dg);
                            dwacc[6 + (2 - sx)] = f4fma(av[o], D0[o + sx], dwacc[6 + (2 - sx)]);
                            dwacc[3 + (2 - sx)] = f4fma(av[o], D1[o + sx], dwacc[3 + (2 - sx)]);
                            dwacc[0 + (2 - sx)] = f4fma(av[o], D2[o + sx], dwacc[0 + (2 - sx)]);
                        }
                        bwd_pixel_finish<T>(cx, g_in, offs[o], dg, yin[o], skv[o], ls1, ls2);
                    }
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) { D0[j] = D1[j]; D1[j] = D2[j]; }
            }
        } else {
            // 2x2 input blocks: block (a, cb) = input rows 2a, 2a+1 x cols 2cb, 2cb+1 receives from dy(a..a+1, cb..cb+1)
            constexpr int NB = 1;
            const int cb0 = it.wg * NB;
            float4 E0[NB + 1], E1[NB + 1];
            load_row<T, DM, NB + 1>(dy, dp, oimg, it.r_a, gm.Ho, cb0, gm.Wo, gm.C, c, E0);
            for (int a = it.r_a; a < it.r_b; ++a) {
                float4 yin[NB][4], av[NB][4], skv[NB][4], rg[NB + 1], ry[NB + 1];
                size_t offs[NB][4];
                load_row_raw<T, DM, NB + 1>(dy, oimg, a + 1, gm.Ho, cb0, gm.Wo, gm.C, c, rg, ry);
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int hi = min(2 * a + (q >> 1), gm.H - 1), wi = min(2 * (cb0 + j) + (q & 1), gm.W - 1);
                        offs[j][q] = (iimg + (size_t)hi * gm.W + wi) * gm.C + c;
                        pixel_loads(offs[j][q], yin[j][q], av[j][q], skv[j][q]);
                    }
                apply_row<DM, NB + 1>(dp, a + 1, gm.Ho, cb0, gm.Wo, rg, ry, E1);
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) av[j][q] = in_val(yin[j][q], av[j][q]);
#pragma unroll
                for (int j = 0; j < NB; ++j) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ph = q >> 1, pw = q & 1;
                        if (2 * a + ph < gm.H && 2 * (cb0 + j) + pw < gm.W) {
                            const float4 a4 = av[j][q];
                            float4 dg;
                            if (ph == 0 && pw == 0) {
                                dg = f4mul(E0[j], sh.w[4][cl]);
                                dwacc[4] = f4fma(a4, E0[j], dwacc[4]);
                            } else if (ph == 0) {
                                dg = f4mul(E0[j], sh.w[5][cl]); dg = f4fma(E0[j + 1], sh.w[3][cl], dg);
                                dwacc[5] = f4fma(a4, E0[j], dwacc[5]); dwacc[3] = f4fma(a4, E0[j + 1], dwacc[3]);
                            } else if (pw == 0) {
                                dg = f4mul(E0[j], sh.w[7][cl]); dg = f4fma(E1[j], sh.w[1][cl], dg);
                                dwacc[7] = f4fma(a4, E0[j], dwacc[7]); dwacc[1] = f4fma(a4, E1[j], dwacc[1]);
                            } else {
                                dg = f4mul(E0[j], sh.w[8][cl]); dg = f4fma(E0[j + 1], sh.w[6][cl], dg);
                                dg = f4fma(E1[j], sh.w[2][cl], dg); dg = f4fma(E1[j + 1], sh.w[0][cl], dg);
                                dwacc[8] = f4fma(a4, E0[j], dwacc[8]); dwacc[6] = f4fma(a4, E0[j + 1], dwacc[6]);
                                dwacc[2] = f4fma(a4, E1[j], dwacc[2]); dwacc[0] = f4fma(a4, E1[j + 1], dwacc[0]);
                            }
                            bwd_pixel_finish<T>(cx, g_in, offs[j][q], dg, yin[j][q], skv[j][q], ls1, ls2);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NB + 1; ++j) E0[j] = E1[j];
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            atomicAdd(&sh.dw[t][cl * 4 + 0], dwacc[t].x); atomicAdd(&sh.dw[t][cl * 4 + 1], dwacc[t].y);
            atomicAdd(&sh.dw[t][cl * 4 + 2], dwacc[t].z); atomicAdd(&sh.dw[t][cl * 4 + 3], dwacc[t].w);
        }
        if (cx.do_stats) {
            // s2 = sum g*xhat = rstd * (sum g*y - mean * sum g)
            const float4 mu = ldg4(bn.mean + c), rs = ldg4(bn.rstd + c);
            atomicAdd(&sh.s1[cl * 4 + 0], ls1.x); atomicAdd(&sh.s1[cl * 4 + 1], ls1.y);
            atomicAdd(&sh.s1[cl * 4 + 2], ls1.z); atomicAdd(&sh.s1[cl * 4 + 3], ls1.w);
            atomicAdd(&sh.s2[cl * 4 + 0], rs.x * (ls2.x - mu.x * ls1.x)); atomicAdd(&sh.s2[cl * 4 + 1], rs.y * (ls2.y - mu.y * ls1.y));
            atomicAdd(&sh.s2[cl * 4 + 2], rs.z * (ls2.z - mu.z * ls1.z)); atomicAdd(&sh.s2[cl * 4 + 3], rs.w * (ls2.w - mu.w * ls1.w));
        }
    }
    __syncthreads();
    if (tid < gm.CB * 4) {
        for (int t = 0; t < 9; ++t) atomicAdd(dw9c + (size_t)t * gm.C + cbase + tid, sh.dw[t][tid]);
        if (cx.do_stats) {
            atomicAdd(bn.s1 + cbase + tid, (double)sh.s1[tid]);
            atomicAdd(bn.s2 + cbase + tid, (double)sh.s2[tid]);
        }
    }
    if (cx.do_stats && grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_bwd_finalize_all(bn, gm.C, gm.count, tid, RNT);
}

// ------------------------------------------------------------------------------------------------
// Backward, second generation (round-2 candidate, selected with B200SP_DW=2; v1 above stays the default until measured).
// Specialised for the shape every MobileNetV2 block has -- dy is a BatchNorm-backward virtual tensor (XM_DY), the conv input is
// act(BN(y_in)) of the SAME tensor whose BatchNorm receives the gradient, activation ReLU / ReLU6, no skip gradient -- and written
// against the source-level profile of v1 (profiles/r1_o_ncu_source_lines.txt): 773 SASS instructions per 2x2 block for ~150
// essential FFMAs, 168 registers (12 warps/SM), and 44 compare-and-swap loops per thread for the shared-memory float atomics.
//   * activation fixed at compile time (2 FMNMX per value instead of the 4-instruction branch-free generic form),
//   * 32-bit element offsets, column offsets and column validity hoisted out of the row loop,
//   * out-of-range dy taps zeroed by a 0/1 multiplier instead of per-value selects,
//   * weight-gradient / statistic partials staged in shared memory with plain vector stores and summed by a few threads
//     (one red.global.add.v4.f32 per 16 bytes), no shared atomics.
template <int ACT> __device__ __forceinline__ float actf2(float z) {
    return ACT == B200SP_ACT_RELU6 ? fminf(fmaxf(z, 0.f), 6.f) : fmaxf(z, 0.f);
}
template <int ACT> __device__ __forceinline__ bool actd2(float z) {            // act'(z) != 0  (open interval, like act_bwd)
    return ACT == B200SP_ACT_RELU6 ? (z > 0.f && z < 6.f) : (z > 0.f);
}
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 ldf4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct Bwd2Shared {
    float4 w[9][MAXCB];
    float4 dw[RNT][9];              // per-thread weight-gradient partials, [thread][tap]: stride 144 B -> conflict-free
    float4 s1[RNT], s2[RNT];
};

// one input pixel: z = BN pre-activation, av = conv input value, g = (dgrad) * act'(z); statistics partials
template <int ACT>
__device__ __forceinline__ float4 bwd2_finish(float4 dg, float4 yin, float4 z, float4& ls1, float4& ls2) {
    dg.x = actd2<ACT>(z.x) ? dg.x : 0.f; dg.y = actd2<ACT>(z.y) ? dg.y : 0.f;
    dg.z = actd2<ACT>(z.z) ? dg.z : 0.f; dg.w = actd2<ACT>(z.w) ? dg.w : 0.f;
    ls1 = f4add(ls1, dg);
    ls2 = f4fma(dg, yin, ls2);
    return dg;
}

// PF (stride 1): the raw loads of the NEXT row step are issued before the current step's arithmetic (register double buffer,
// 2 CTAs per SM instead of 3) -- without it a thread's loads and its ~200 dependent instructions per row strictly alternate.
template <int S, int ACT, bool PF = false>
__global__ void __launch_bounds__(RNT, PF ? (S == 2 ? 3 : 2) : (S == 2 ? 4 : DWR_S1_OCC)) dwr_bwd2_kernel(const b200sp_vtensor dy, const float* __restrict__ w9c,
                                                          float* __restrict__ g_in, float* __restrict__ dw9c,
                                                          const b200sp_bnbwd bn, const RGeom gm) {
    extern __shared__ __align__(16) unsigned char bwd2_smem[];
    Bwd2Shared& sh = *reinterpret_cast<Bwd2Shared*>(bwd2_smem);
    pdl_trigger();          // programmatic dependent launch (common.cuh): block scheduling overlapped the predecessor's tail
    pdl_wait();
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int i = tid; i < 9 * gm.CB; i += RNT) sh.w[i / gm.CB][i % gm.CB] = ldg4(w9c + (size_t)(i / gm.CB) * gm.C + cbase + (i % gm.CB) * 4);
    __syncthreads();

    const Item it = get_item(gm, sl, S == 1 ? gm.H : gm.Ho);
    float4 ls1 = f4zero(), ls2 = f4zero();
    float4 dwacc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dwacc[t] = f4zero();
    if (it.ok) {
        const float4 cA = ldg4(dy.p0 + c), cB = ldg4(dy.p1 + c), cC = ldg4(dy.p2 + c);      // dy = cA*g + cB*y + cC
        const float4 sc = ldg4(bn.scale + c), shf = ldg4(bn.shift + c);
        const float* __restrict__ gq = reinterpret_cast<const float*>(dy.x);
        const float* __restrict__ yq = reinterpret_cast<const float*>(dy.x2);
        const float* __restrict__ yb = reinterpret_cast<const float*>(bn.y);
        const int C = gm.C;
        auto dyv = [&](int off, float m) {           // transformed dy at element offset `off`, times the 0/1 validity factor
            const float4 g = ldf4(gq + off), y = ldf4(yq + off);
            return f4scale(f4fma(cA, g, f4fma(cB, y, cC)), m);
        };
        if (S == 2) {
            // item = block column cb (input columns 2cb, 2cb+1), block rows [r_a, r_b): input rows 2a, 2a+1 receive dy(a..a+1, cb..cb+1)
            const int cb = it.wg;
            const bool c1 = cb + 1 < gm.Wo, w1 = 2 * cb + 1 < gm.W;
            const float m1 = c1 ? 1.f : 0.f;
            const int dcol = c1 ? C : 0;                                   // dy column cb+1 (clamped onto cb when outside)
            const int drow = gm.Wo * C, xrow = gm.W * C, xcol = w1 ? C : 0;
            int od = ((it.b * gm.Ho + it.r_a) * gm.Wo + cb) * C + c;       // dy(a, cb)
            int ox = ((it.b * gm.H + 2 * it.r_a) * gm.W + 2 * cb) * C + c; // input (2a, 2cb)
            float4 E00 = dyv(od, 1.f), E01 = dyv(od + dcol, m1);
            float4 pf[8];                                  // PF: raw g10 y10 g11 y11 yi0..yi3 of the step about to run
            auto fetch = [&](int a, int od_, int ox_, float4 (&q)[8]) {
                const int odn = od_ + (a + 1 < gm.Ho ? drow : 0), oxh = ox_ + (2 * a + 1 < gm.H ? xrow : 0);
                q[0] = ldf4(gq + odn); q[1] = ldf4(yq + odn); q[2] = ldf4(gq + odn + dcol); q[3] = ldf4(yq + odn + dcol);
                q[4] = ldf4(yb + ox_); q[5] = ldf4(yb + ox_ + xcol); q[6] = ldf4(yb + oxh); q[7] = ldf4(yb + oxh + xcol);
            };
            if (PF) fetch(it.r_a, od, ox, pf);
            for (int a = it.r_a; a < it.r_b; ++a) {
                const bool r1 = a + 1 < gm.Ho, h1 = 2 * a + 1 < gm.H;
                const float mr = r1 ? 1.f : 0.f;
                const int oxh = ox + (h1 ? xrow : 0);
                // all loads of the step first (PF: they were issued one step ago; the next step's go out now)
                float4 cur[8];
                if (PF) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) cur[j] = pf[j];
                    if (a + 1 < it.r_b) fetch(a + 1, od + drow, ox + 2 * xrow, pf);
                } else {
                    fetch(a, od, ox, cur);
                }
                const float4 g10 = cur[0], y10 = cur[1], g11 = cur[2], y11 = cur[3];
                const float4 yi0 = cur[4], yi1 = cur[5], yi2 = cur[6], yi3 = cur[7];
                const float4 E10 = f4scale(f4fma(cA, g10, f4fma(cB, y10, cC)), mr);
                const float4 E11 = f4scale(f4fma(cA, g11, f4fma(cB, y11, cC)), mr * m1);
                {   // (ph, pw) = (0, 0): tap 4
                    const float4 z = f4fma(yi0, sc, shf);
                    const float4 av = make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w));
                    float4 dg = f4mul(E00, sh.w[4][cl]);
                    dwacc[4] = f4fma(av, E00, dwacc[4]);
                    Vec4<float>::st(g_in + ox, bwd2_finish<ACT>(dg, yi0, z, ls1, ls2));
                }
                if (w1) {   // (0, 1): taps 5, 3
                    const float4 z = f4fma(yi1, sc, shf);
                    const float4 av = make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w));
                    float4 dg = f4mul(E00, sh.w[5][cl]); dg = f4fma(E01, sh.w[3][cl], dg);
                    dwacc[5] = f4fma(av, E00, dwacc[5]); dwacc[3] = f4fma(av, E01, dwacc[3]);
                    Vec4<float>::st(g_in + ox + xcol, bwd2_finish<ACT>(dg, yi1, z, ls1, ls2));
                }
                if (h1) {   // (1, 0): taps 7, 1
                    const float4 z = f4fma(yi2, sc, shf);
                    const float4 av = make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w));
                    float4 dg = f4mul(E00, sh.w[7][cl]); dg = f4fma(E10, sh.w[1][cl], dg);
                    dwacc[7] = f4fma(av, E00, dwacc[7]); dwacc[1] = f4fma(av, E10, dwacc[1]);
                    Vec4<float>::st(g_in + oxh, bwd2_finish<ACT>(dg, yi2, z, ls1, ls2));
                }
                if (h1 && w1) {   // (1, 1): taps 8, 6, 2, 0
                    const float4 z = f4fma(yi3, sc, shf);
                    const float4 av = make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w));
                    float4 dg = f4mul(E00, sh.w[8][cl]); dg = f4fma(E01, sh.w[6][cl], dg);
                    dg = f4fma(E10, sh.w[2][cl], dg); dg = f4fma(E11, sh.w[0][cl], dg);
                    dwacc[8] = f4fma(av, E00, dwacc[8]); dwacc[6] = f4fma(av, E01, dwacc[6]);
                    dwacc[2] = f4fma(av, E10, dwacc[2]); dwacc[0] = f4fma(av, E11, dwacc[0]);
                    Vec4<float>::st(g_in + oxh + xcol, bwd2_finish<ACT>(dg, yi3, z, ls1, ls2));
                }
                E00 = E10; E01 = E11;
                od += drow; ox += 2 * xrow;
            }
        } else {
            // item = input columns wi0, wi0+1, input rows [r_a, r_b); dy columns wi0-1 .. wi0+2 (NC = 4), rows hi-1 .. hi+1 rolling
            const int wi0 = it.wg * 2;
            const bool p1 = wi0 + 1 < gm.W;
            int dco[4];
            float mc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wi0 - 1 + j;
                mc[j] = (col >= 0 && col < gm.Wo) ? 1.f : 0.f;
                dco[j] = min(max(col, 0), gm.Wo - 1) * C;
            }
            const int drow = gm.Wo * C;
            const int dbase = (it.b * gm.Ho) * gm.Wo * C + c;              // dy(b, 0, 0)
            auto dyrow = [&](int row, float4 (&D)[4]) {
                const float mr = (row >= 0 && row < gm.Ho) ? 1.f : 0.f;
                const int o = dbase + min(max(row, 0), gm.Ho - 1) * drow;
#pragma unroll
                for (int j = 0; j < 4; ++j) D[j] = dyv(o + dco[j], mr * mc[j]);
            };
            float4 D0[4], D1[4], D2[4];
            dyrow(it.r_a - 1, D0);
            dyrow(it.r_a, D1);
            int ox = ((it.b * gm.H + it.r_a) * gm.W + wi0) * C + c;
            const int xcol = p1 ? C : 0, xrow = gm.W * C;
            float4 pg[4], py[4], pyi0, pyi1;              // PF: raw values of the step about to run
            if (PF) {
                const int o = dbase + min(it.r_a + 1, gm.Ho - 1) * drow;
#pragma unroll
                for (int j = 0; j < 4; ++j) { pg[j] = ldf4(gq + o + dco[j]); py[j] = ldf4(yq + o + dco[j]); }
                pyi0 = ldf4(yb + ox); pyi1 = ldf4(yb + ox + xcol);
            }
            for (int hi = it.r_a; hi < it.r_b; ++hi) {
                const int rn = hi + 1;
                const float mr = rn < gm.Ho ? 1.f : 0.f;
                const int o = dbase + min(rn, gm.Ho - 1) * drow;
                float4 g[4], y[4];
                float4 yi0, yi1;
                if (PF) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { g[j] = pg[j]; y[j] = py[j]; }
                    yi0 = pyi0; yi1 = pyi1;
                    if (hi + 1 < it.r_b) {                  // next step: dy row hi + 2 (clamped), input row hi + 1
                        const int o2 = dbase + min(rn + 1, gm.Ho - 1) * drow;
#pragma unroll
                        for (int j = 0; j < 4; ++j) { pg[j] = ldf4(gq + o2 + dco[j]); py[j] = ldf4(yq + o2 + dco[j]); }
                        pyi0 = ldf4(yb + ox + xrow); pyi1 = ldf4(yb + ox + xrow + xcol);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { g[j] = ldf4(gq + o + dco[j]); y[j] = ldf4(yq + o + dco[j]); }
                    yi0 = ldf4(yb + ox); yi1 = ldf4(yb + ox + xcol);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) D2[j] = f4scale(f4fma(cA, g[j], f4fma(cB, y[j], cC)), mr * mc[j]);
#pragma unroll
                for (int o2 = 0; o2 < 2; ++o2) {
                    if (o2 == 0 || p1) {
                        const float4 yi = o2 == 0 ? yi0 : yi1;
                        const float4 z = f4fma(yi, sc, shf);
                        const float4 av = make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w));
                        float4 dg = f4zero();
                        // output (hi+dh, wi+dw) used tap (kh,kw) = (1-dh, 1-dw);  D<r>[o2+sx] holds dh = r-1, dw = sx-1
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            dg = f4fma(D0[o2 + sx], sh.w[6 + (2 - sx)][cl], dg);
                            dg = f4fma(D1[o2 + sx], sh.w[3 + (2 - sx)][cl], dg);
                            dg = f4fma(D2[o2 + sx], sh.w[0 + (2 - sx)][cl], dg);
                            dwacc[6 + (2 - sx)] = f4fma(av, D0[o2 + sx], dwacc[6 + (2 - sx)]);
                            dwacc[3 + (2 - sx)] = f4fma(av, D1[o2 + sx], dwacc[3 + (2 - sx)]);
                            dwacc[0 + (2 - sx)] = f4fma(av, D2[o2 + sx], dwacc[0 + (2 - sx)]);
                        }
                        Vec4<float>::st(g_in + ox + (o2 ? xcol : 0), bwd2_finish<ACT>(dg, yi, z, ls1, ls2));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { D0[j] = D1[j]; D1[j] = D2[j]; }
                ox += xrow;
            }
        }
        // s2 = sum g*xhat = rstd * (sum g*y - mean * sum g)
        const float4 mu = ldg4(bn.mean + c), rs = ldg4(bn.rstd + c);
        ls2 = make_float4(rs.x * (ls2.x - mu.x * ls1.x), rs.y * (ls2.y - mu.y * ls1.y), rs.z * (ls2.z - mu.z * ls1.z), rs.w * (ls2.w - mu.w * ls1.w));
    }
    // ---- CTA-level reduction of the partials: plain vector stores, then (tap, channel quad) owners sum over the items ----
#pragma unroll
    for (int t = 0; t < 9; ++t) sh.dw[tid][t] = dwacc[t];
    sh.s1[tid] = ls1;
    sh.s2[tid] = ls2;
    __syncthreads();
    const int nsl = RNT / gm.CB;                     // items per CTA (threads beyond nsl*CB hold zeros and are not read)
    for (int i = tid; i < 9 * gm.CB; i += RNT) {
        const int t = i / gm.CB, q = i - t * gm.CB;
        float4 a = f4zero();
        for (int s2i = 0; s2i < nsl; ++s2i) a = f4add(a, sh.dw[s2i * gm.CB + q][t]);
        float* dst = dw9c + (size_t)t * gm.C + cbase + q * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    }
    if (tid < gm.CB * 4) {
        const int q = tid >> 2, k = tid & 3;
        float a = 0.f, b = 0.f;
        for (int s2i = 0; s2i < nsl; ++s2i) {
            a += reinterpret_cast<const float*>(&sh.s1[s2i * gm.CB + q])[k];
            b += reinterpret_cast<const float*>(&sh.s2[s2i * gm.CB + q])[k];
        }
        atomicAdd(bn.s1 + cbase + tid, (double)a);
        atomicAdd(bn.s2 + cbase + tid, (double)b);
    }
    if (grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_bwd_finalize_all(bn, gm.C, gm.count, tid, RNT);
}

// Forward, second generation (same recipe as dwr_bwd2_kernel; B200SP_DW=2).  Input = act(BN(y_prev)) with the activation fixed
// at compile time, zero padding applied as a 0/1 multiplier, 32-bit offsets with the column part hoisted out of the row loop,
// BatchNorm statistic partials staged in shared memory (no shared atomics).  v1: 539 / 520 SASS instructions per row step
// (stride 1 / 2) for 144 / 72 essential FFMAs.
struct Fwd2Shared {
    float4 w[9][MAXCB];
    float4 s1[RNT], s2[RNT];
};

template <int S, int ACT, bool PF = false>
__global__ void __launch_bounds__(RNT, PF ? (S == 2 ? 3 : 2) : (S == 2 ? 4 : DWR_S1_OCC)) dwr_fwd2_kernel(const b200sp_vtensor x, const float* __restrict__ w9c,
                                                                       float* __restrict__ y, const b200sp_bnfwd bn, const RGeom gm) {
    constexpr int OW = S == 1 ? 4 : 2;
    constexpr int NC = (OW - 1) * S + 3;
    __shared__ Fwd2Shared sh;
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int cl = tid % gm.CB, sl = tid / gm.CB;
    const int cbase = blockIdx.y * gm.CB * 4;
    const int c = cbase + cl * 4;
    for (int i = tid; i < 9 * gm.CB; i += RNT) sh.w[i / gm.CB][i % gm.CB] = ldg4(w9c + (size_t)(i / gm.CB) * gm.C + cbase + (i % gm.CB) * 4);
    __syncthreads();

    const Item it = get_item(gm, sl, gm.Ho);
    float4 lsum = f4zero(), lsq = f4zero();
    if (it.ok) {
        const float4 sc = ldg4(x.p0 + c), shf = ldg4(x.p1 + c);
        const float* __restrict__ xq = reinterpret_cast<const float*>(x.x);
        const int C = gm.C;
        const int wo0 = it.wg * OW, wi0 = wo0 * S - 1;
        int xco[NC];
        float mc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int col = wi0 + j;
            mc[j] = (col >= 0 && col < gm.W) ? 1.f : 0.f;
            xco[j] = min(max(col, 0), gm.W - 1) * C;
        }
        const int xrow = gm.W * C;
        const int xbase = it.b * gm.H * xrow + c;
        auto loadrow = [&](int row, float4 (&r)[NC]) {
            const float mr = (row >= 0 && row < gm.H) ? 1.f : 0.f;
            const int o = xbase + min(max(row, 0), gm.H - 1) * xrow;
            float4 raw[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) raw[j] = ldf4(xq + o + xco[j]);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 z = f4fma(raw[j], sc, shf);
                r[j] = f4scale(make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w)), mr * mc[j]);
            }
        };
        auto rawrow = [&](int row, float4 (&raw)[NC]) {        // PF: issue only; validity is applied when the row is consumed
            const int o = xbase + min(max(row, 0), gm.H - 1) * xrow;
#pragma unroll
            for (int j = 0; j < NC; ++j) raw[j] = ldf4(xq + o + xco[j]);
        };
        auto finrow = [&](int row, const float4 (&raw)[NC], float4 (&r)[NC]) {
            const float mr = (row >= 0 && row < gm.H) ? 1.f : 0.f;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 z = f4fma(raw[j], sc, shf);
                r[j] = f4scale(make_float4(actf2<ACT>(z.x), actf2<ACT>(z.y), actf2<ACT>(z.z), actf2<ACT>(z.w)), mr * mc[j]);
            }
        };
        float4 r0[NC], r1[NC], r2[NC];
        float4 nx[NC], nx2[S == 2 ? NC : 1];
        if (S == 1) { loadrow(it.r_a - 1, r0); loadrow(it.r_a, r1); if (PF) rawrow(it.r_a + 1, nx); }
        else        { loadrow(2 * it.r_a - 1, r0); if (PF) { rawrow(2 * it.r_a, nx); rawrow(2 * it.r_a + 1, reinterpret_cast<float4(&)[NC]>(nx2)); } }
        int oy = ((it.b * gm.Ho + it.r_a) * gm.Wo + wo0) * C + c;
        const int yrow = gm.Wo * C;
        for (int ho = it.r_a; ho < it.r_b; ++ho) {
            if (S == 1 && PF) { finrow(ho + 1, nx, r2); if (ho + 1 < it.r_b) rawrow(ho + 2, nx); }
            else if (S == 1) { loadrow(ho + 1, r2); }
            else if (PF) {
                finrow(2 * ho, nx, r1); finrow(2 * ho + 1, reinterpret_cast<float4(&)[NC]>(nx2), r2);
                if (ho + 1 < it.r_b) { rawrow(2 * ho + 2, nx); rawrow(2 * ho + 3, reinterpret_cast<float4(&)[NC]>(nx2)); }
            }
            else        { loadrow(2 * ho, r1); loadrow(2 * ho + 1, r2); }
#pragma unroll
            for (int o = 0; o < OW; ++o) {
                float4 acc = f4mul(r0[o * S], sh.w[0][cl]);
                acc = f4fma(r0[o * S + 1], sh.w[1][cl], acc); acc = f4fma(r0[o * S + 2], sh.w[2][cl], acc);
                acc = f4fma(r1[o * S], sh.w[3][cl], acc); acc = f4fma(r1[o * S + 1], sh.w[4][cl], acc); acc = f4fma(r1[o * S + 2], sh.w[5][cl], acc);
                acc = f4fma(r2[o * S], sh.w[6][cl], acc); acc = f4fma(r2[o * S + 1], sh.w[7][cl], acc); acc = f4fma(r2[o * S + 2], sh.w[8][cl], acc);
                if (wo0 + o < gm.Wo) {
                    Vec4<float>::st(y + oy + o * C, acc);
                    lsum = f4add(lsum, acc);
                    lsq = f4fma(acc, acc, lsq);
                }
            }
            if (S == 1) {
#pragma unroll
                for (int j = 0; j < NC; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < NC; ++j) r0[j] = r2[j];
            }
            oy += yrow;
        }
    }
    sh.s1[tid] = lsum;
    sh.s2[tid] = lsq;
    __syncthreads();
    if (tid < gm.CB * 4) {
        const int nsl = RNT / gm.CB, q = tid >> 2, k = tid & 3;
        float a = 0.f, b = 0.f;
        for (int s2i = 0; s2i < nsl; ++s2i) {
            a += reinterpret_cast<const float*>(&sh.s1[s2i * gm.CB + q])[k];
            b += reinterpret_cast<const float*>(&sh.s2[s2i * gm.CB + q])[k];
        }
        atomicAdd(bn.sum + cbase + tid, (double)a);
        atomicAdd(bn.sumsq + cbase + tid, (double)b);
    }
    if (grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_fwd_finalize_all(bn, gm.C, gm.count, tid, RNT);
}

// rows: number of rolled rows; cols: number of column groups; count: BatchNorm population
int roll_geom(RGeom& gm, dim3& grid, int B, int H, int W, int C, int stride, int rows, int ncolgroups, double count) {
    if (C % 4 || (stride != 1 && stride != 2)) return B200SP_EINVAL;
    gm.B = B; gm.H = H; gm.W = W; gm.C = C;
    gm.Ho = (H - 1) / stride + 1; gm.Wo = (W - 1) / stride + 1;
    const int C4 = C / 4;
    int cb = 1;
    for (int d = 1; d <= MAXCB && d <= C4; ++d) if (C4 % d == 0) cb = d;
    gm.CB = cb; gm.SPC = RNT / cb;
    gm.nWG = ncolgroups;
    // split the rolled dimension until ~12 warps/SM x 2 waves of threads exist (segments of >= 8 rows: halo overhead <= 25 %).
    // Finer segments for the small maps (B200SP_DW_SMALLSEG=1: >= 3 rows, 3-4x more CTAs) were measured and are SLOWER
    // (dw_bwd 1214 -> 1323 us per step, job r2v; whole step 5.40 -> 5.49 / 5.81 / 6.37 ms with segments of >= 4 / 2 / 1 rows, job r3u):
    // those launches are bound by their per-CTA reductions (same-address atomics), not by parallelism.  Whole-image "slab" kernels
    // for the 7x7 / 14x14 maps (all loads issued at once into shared memory, one CTA per 1-4 images and 64 channels) were written
    // and measured too: 5.43 vs 5.40 ms in-graph -- no gain, not kept (job r3v).
    const long long base = (long long)B * ncolgroups * C4;
    const long long target = (long long)NUM_SMS * 3072;
    int nseg = (int)((target + base - 1) / base);
    static int small_seg = -1;       // B200SP_DW_SMALLSEG=<n>: maps below 32 rows are cut into segments of >= n rows (0: off)
    if (small_seg < 0) { const char* e = getenv("B200SP_DW_SMALLSEG"); small_seg = e ? atoi(e) : 0; if (small_seg < 0) small_seg = 0; }
    const int maxseg = rows >= 32 ? rows / 8 : (small_seg ? (rows >= small_seg ? rows / small_seg : 1) : (rows >= 16 ? rows / 8 : 1));
    if (nseg > maxseg) nseg = maxseg;
    if (nseg < 1) nseg = 1;
    gm.SEG = (rows + nseg - 1) / nseg;
    gm.nSeg = (rows + gm.SEG - 1) / gm.SEG;
    gm.nitems = (long long)B * gm.nSeg * ncolgroups;
    gm.count = count;
    grid = dim3((unsigned)((gm.nitems + gm.SPC - 1) / gm.SPC), C4 / cb, 1);
    return 0;
}

inline bool use_roll() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_DW"); v = (e && e[0] == 'l') ? 0 : 1; }
    return v == 1;
}

inline int dw_prefetch() {         // B200SP_DW_PF: 0 none | 1 stride-1 kernels (default) | 2 stride-2 kernels as well
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_DW_PF"); v = !e ? 1 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)); }
    return v;
}

inline bool use_bwd2() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_DW"); v = (e && e[0] == '1') ? 0 : 1; }      // second-generation kernels by default (validated on B200 in round 2: -0.4 ms/step); B200SP_DW=1 selects the first generation
    return v == 1;
}

template <typename T>
int launch_fwd(const b200sp_vtensor* x, const float* w9c, void* y, const b200sp_bnfwd* bn, int B, int H, int W, int C, int stride, cudaStream_t st) {
    RGeom gm; dim3 grid;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int OW = stride == 1 ? 4 : 2;
    if (int rc = roll_geom(gm, grid, B, H, W, C, stride, Ho, (Wo + OW - 1) / OW, (double)B * Ho * Wo)) return rc;
    b200sp_bnfwd b = {};
    if (bn) b = *bn;
    const bool plain = x->mode == B200SP_VT_PLAIN;
    if (use_bwd2() && sizeof(T) == 4 && bn && x->mode == B200SP_VT_BNACT && (x->act == B200SP_ACT_RELU6 || x->act == B200SP_ACT_RELU) &&
        (long long)B * H * W * C < (1ll << 31)) {
        const bool r6 = x->act == B200SP_ACT_RELU6;
        if (stride == 1 && dw_prefetch()) {
            if (r6) b200sp_launch_pdl(dwr_fwd2_kernel<1, B200SP_ACT_RELU6, true>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
            else    b200sp_launch_pdl(dwr_fwd2_kernel<1, B200SP_ACT_RELU, true>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
        } else if (stride == 1) {
            if (r6) b200sp_launch_pdl(dwr_fwd2_kernel<1, B200SP_ACT_RELU6>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
            else    b200sp_launch_pdl(dwr_fwd2_kernel<1, B200SP_ACT_RELU>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
        } else if (dw_prefetch() == 2) {
            if (r6) b200sp_launch_pdl(dwr_fwd2_kernel<2, B200SP_ACT_RELU6, true>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
            else    b200sp_launch_pdl(dwr_fwd2_kernel<2, B200SP_ACT_RELU, true>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
        } else {
            if (r6) b200sp_launch_pdl(dwr_fwd2_kernel<2, B200SP_ACT_RELU6>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
            else    b200sp_launch_pdl(dwr_fwd2_kernel<2, B200SP_ACT_RELU>, grid, dim3(RNT), 0, st, *x, w9c, (float*)y, b, gm);
        }
        B200SP_COUNT_LAUNCH();
        B200SP_RETURN_LAST();
    }
    if (stride == 1) {
        if (plain) dwr_fwd_kernel<T, 1, XM_PLAIN><<<grid, RNT, 0, st>>>(*x, w9c, (T*)y, b, bn != nullptr, gm);
        else       dwr_fwd_kernel<T, 1, XM_BNACT><<<grid, RNT, 0, st>>>(*x, w9c, (T*)y, b, bn != nullptr, gm);
    } else {
        if (plain) dwr_fwd_kernel<T, 2, XM_PLAIN><<<grid, RNT, 0, st>>>(*x, w9c, (T*)y, b, bn != nullptr, gm);
        else       dwr_fwd_kernel<T, 2, XM_BNACT><<<grid, RNT, 0, st>>>(*x, w9c, (T*)y, b, bn != nullptr, gm);
    }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

template <typename T>
int launch_bwd(const b200sp_vtensor* dy, const b200sp_vtensor* x, const float* w9c, const void* skip, void* g_in, float* dw9c,
               const b200sp_bnbwd* bn, int B, int H, int W, int C, int stride, cudaStream_t st) {
    RGeom gm; dim3 grid;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int rows = stride == 1 ? H : Ho;
    const int ncg = stride == 1 ? (W + 1) / 2 : Wo;
    if (int rc = roll_geom(gm, grid, B, H, W, C, stride, rows, ncg, (double)B * H * W)) return rc;
    b200sp_bnbwd b = {};
    if (bn) b = *bn;
    if (use_bwd2() && sizeof(T) == 4 && bn && dy->mode == B200SP_VT_DY && x->mode == B200SP_VT_BNACT && x->x == bn->y && !skip && g_in &&
        bn->scale && bn->s1 && x->act == bn->act && (bn->act == B200SP_ACT_RELU6 || bn->act == B200SP_ACT_RELU) &&
        x->p0 == bn->scale && x->p1 == bn->shift &&
        (long long)B * H * W * C < (1ll << 31) && (long long)B * Ho * Wo * C < (1ll << 31)) {
        static bool attr_set = false;
        const int smem = (int)sizeof(Bwd2Shared);
        if (!attr_set) {
            cudaFuncSetAttribute(dwr_bwd2_kernel<1, B200SP_ACT_RELU6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<2, B200SP_ACT_RELU6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<1, B200SP_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<2, B200SP_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<1, B200SP_ACT_RELU6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<1, B200SP_ACT_RELU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<2, B200SP_ACT_RELU6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(dwr_bwd2_kernel<2, B200SP_ACT_RELU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set = true;
        }
        const bool r6 = bn->act == B200SP_ACT_RELU6;
        if (stride == 1 && dw_prefetch()) {
            if (r6) b200sp_launch_pdl(dwr_bwd2_kernel<1, B200SP_ACT_RELU6, true>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
            else    b200sp_launch_pdl(dwr_bwd2_kernel<1, B200SP_ACT_RELU, true>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
        } else if (stride == 1) {
            if (r6) b200sp_launch_pdl(dwr_bwd2_kernel<1, B200SP_ACT_RELU6>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
            else    b200sp_launch_pdl(dwr_bwd2_kernel<1, B200SP_ACT_RELU>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
        } else if (dw_prefetch() == 2) {
            if (r6) b200sp_launch_pdl(dwr_bwd2_kernel<2, B200SP_ACT_RELU6, true>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
            else    b200sp_launch_pdl(dwr_bwd2_kernel<2, B200SP_ACT_RELU, true>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
        } else {
            if (r6) b200sp_launch_pdl(dwr_bwd2_kernel<2, B200SP_ACT_RELU6>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
            else    b200sp_launch_pdl(dwr_bwd2_kernel<2, B200SP_ACT_RELU>, grid, dim3(RNT), smem, st, *dy, w9c, (float*)g_in, dw9c, b, gm);
        }
        B200SP_COUNT_LAUNCH();
        B200SP_RETURN_LAST();
    }
#define DWR_BWD(S_, DM_, XM_) dwr_bwd_kernel<T, S_, DM_, XM_><<<grid, RNT, 0, st>>>(*dy, *x, w9c, (const T*)skip, (T*)g_in, dw9c, b, bn != nullptr, gm)
    const bool ddy = dy->mode == B200SP_VT_DY, xbn = x->mode == B200SP_VT_BNACT;
    if (dy->mode == B200SP_VT_BNACT) return B200SP_EINVAL;
    if (stride == 1) {
        if (ddy && xbn) DWR_BWD(1, XM_DY, XM_BNACT); else if (ddy) DWR_BWD(1, XM_DY, XM_PLAIN);
        else if (xbn) DWR_BWD(1, XM_PLAIN, XM_BNACT); else DWR_BWD(1, XM_PLAIN, XM_PLAIN);
    } else {
        if (ddy && xbn) DWR_BWD(2, XM_DY, XM_BNACT); else if (ddy) DWR_BWD(2, XM_DY, XM_PLAIN);
        else if (xbn) DWR_BWD(2, XM_PLAIN, XM_BNACT); else DWR_BWD(2, XM_PLAIN, XM_PLAIN);
    }
#undef DWR_BWD
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

}  // namespace

extern "C" int b200sp_dw_fwd(const b200sp_vtensor* x, const float* w9c, void* y, const b200sp_bnfwd* bn,
                             int B, int H, int W, int C, int stride, int dtype, void* stream) {
    if (!x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    if (!use_roll()) return dw_fwd_legacy(x, w9c, y, bn, B, H, W, C, stride, dtype, stream);
    if (dtype == B200SP_F32) return launch_fwd<float>(x, w9c, y, bn, B, H, W, C, stride, (cudaStream_t)stream);
    if (dtype == B200SP_BF16) return launch_fwd<bf16>(x, w9c, y, bn, B, H, W, C, stride, (cudaStream_t)stream);
    return B200SP_ENOSYS;
}

extern "C" int b200sp_dw_bwd(const b200sp_vtensor* dy, const b200sp_vtensor* x, const float* w9c, const void* skip,
                             void* g_in, float* dw9c, const b200sp_bnbwd* bn,
                             int B, int H, int W, int C, int stride, int dtype, void* stream) {
    if (!dy || !x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    if (!use_roll()) return dw_bwd_legacy(dy, x, w9c, skip, g_in, dw9c, bn, B, H, W, C, stride, dtype, stream);
    if (dtype == B200SP_F32) return launch_bwd<float>(dy, x, w9c, skip, g_in, dw9c, bn, B, H, W, C, stride, (cudaStream_t)stream);
    if (dtype == B200SP_BF16) return launch_bwd<bf16>(dy, x, w9c, skip, g_in, dw9c, bn, B, H, W, C, stride, (cudaStream_t)stream);
    return B200SP_ENOSYS;
}
