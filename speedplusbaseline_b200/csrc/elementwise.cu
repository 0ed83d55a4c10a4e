// elementwise.cu -- HBM-bound vectorised passes: BatchNorm apply (+residual), standalone BN
// finalisers/reductions, RouterV2 reorg+concat, the KRN 7x7 head (as an FC) and its loss.
#include "common.cuh"

namespace {

constexpr int EW_NT = 256;

// out = act(y*scale + shift) (+ residual); mobilenetv2.py:61-62 (block output) and eval-time affine
template <typename T>
__global__ void __launch_bounds__(EW_NT) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, const T* __restrict__ res,
                                                         int act, T* __restrict__ out, long long n4, int C4) {
    pdl_trigger();
    pdl_wait();
    // four independent 16-byte pieces per thread and iteration (all loads issued before any arithmetic), 32-bit index arithmetic:
    // one piece per iteration with a 64-bit modulo ran at 1.8 TB/s
    if (n4 < (1ll << 30)) {
        const unsigned stride = gridDim.x * EW_NT, n = (unsigned)n4, uc4 = (unsigned)C4;
        unsigned i = blockIdx.x * EW_NT + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            float4 v[4], r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = Vec4<T>::ld(y + (size_t)(i + u * stride) * 4);
            if (res) {
#pragma unroll
                for (int u = 0; u < 4; ++u) r[u] = Vec4<T>::ld(res + (size_t)(i + u * stride) * 4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = (int)((i + u * stride) % uc4) * 4;
                const float4 a = ldg4(scale + c), b = ldg4(shift + c);
                float4 o = make_float4(act_fwd(fmaf(v[u].x, a.x, b.x), act), act_fwd(fmaf(v[u].y, a.y, b.y), act),
                                       act_fwd(fmaf(v[u].z, a.z, b.z), act), act_fwd(fmaf(v[u].w, a.w, b.w), act));
                if (res) { o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w; }
                Vec4<T>::st(out + (size_t)(i + u * stride) * 4, o);
            }
        }
        for (; i < n; i += stride) {
            const int c = (int)(i % uc4) * 4;
            float4 v = Vec4<T>::ld(y + (size_t)i * 4);
            const float4 a = ldg4(scale + c), b = ldg4(shift + c);
            v = make_float4(act_fwd(fmaf(v.x, a.x, b.x), act), act_fwd(fmaf(v.y, a.y, b.y), act),
                            act_fwd(fmaf(v.z, a.z, b.z), act), act_fwd(fmaf(v.w, a.w, b.w), act));
            if (res) { const float4 rr = Vec4<T>::ld(res + (size_t)i * 4); v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
            Vec4<T>::st(out + (size_t)i * 4, v);
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * EW_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * EW_NT) {
        const int c = (int)(i % C4) * 4;
        float4 v = Vec4<T>::ld(y + i * 4);
        const float4 a = ldg4(scale + c), b = ldg4(shift + c);
        v = make_float4(act_fwd(fmaf(v.x, a.x, b.x), act), act_fwd(fmaf(v.y, a.y, b.y), act),
                        act_fwd(fmaf(v.z, a.z, b.z), act), act_fwd(fmaf(v.w, a.w, b.w), act));
        if (res) { const float4 r = Vec4<T>::ld(res + i * 4); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        Vec4<T>::st(out + i * 4, v);
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = gamma[i] / sqrtf(rv[i] + eps);
        scale[i] = s;
        shift[i] = beta[i] - rm[i] * s;
    }
}

__global__ void bn_fwd_finalize_kernel(const b200sp_bnfwd bn, int C, double count) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) bn_fwd_finalize_channel(bn, c, count);
}
__global__ void bn_bwd_finalize_kernel(const b200sp_bnbwd bn, int C, double count) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) bn_bwd_finalize_channel(bn, c, count);
}

// s1 = sum g, s2 = sum g*xhat over rows of a materialised [M,C] gradient; CTA = 64 channels x 4 row lanes
template <typename T>
__global__ void __launch_bounds__(EW_NT) bn_bwd_reduce_kernel(const T* __restrict__ g, const b200sp_bnbwd bn, long long M, int C) {
    __shared__ float s1s[4][64], s2s[4][64];
    const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int c = blockIdx.y * 64 + cl;
    float a = 0.f, b = 0.f;
    if (c < C) {
        const float mu = bn.mean[c], rs = bn.rstd[c];
        const T* y = reinterpret_cast<const T*>(bn.y);
        for (long long r = (long long)blockIdx.x * 4 + rl; r < M; r += (long long)gridDim.x * 4) {
            const float gv = Vec4<T>::ld1(g + r * C + c), yv = Vec4<T>::ld1(y + r * C + c);
            a += gv;
            b = fmaf(gv, (yv - mu) * rs, b);
        }
    }
    s1s[rl][cl] = a; s2s[rl][cl] = b;
    __syncthreads();
    if (threadIdx.x < 64 && c < C) {
        atomicAdd(bn.s1 + c, (double)s1s[0][cl] + (double)s1s[1][cl] + (double)s1s[2][cl] + (double)s1s[3][cl]);
        atomicAdd(bn.s2 + c, (double)s2s[0][cl] + (double)s2s[1][cl] + (double)s2s[2][cl] + (double)s2s[3][cl]);
    }
    if (grid_last_cta(bn.ticket, gridDim.x * gridDim.y))
        bn_bwd_finalize_all(bn, C, (double)M, threadIdx.x, EW_NT);
}

// column sums of a virtual [M,N] tensor (bias gradients): out[n] += sum_m dy[m,n]
template <typename T>
__global__ void __launch_bounds__(EW_NT) colsum_kernel(const b200sp_vtensor dy, float* __restrict__ out, long long M, int N) {
    __shared__ float s[4][64];
    const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int c = blockIdx.y * 64 + cl;
    float a = 0.f;
    if (c < N) {
        float p0 = 1.f, p1 = 0.f, p2 = 0.f;
        if (dy.mode == B200SP_VT_DY) { p0 = dy.p0[c]; p1 = dy.p1[c]; p2 = dy.p2[c]; }
        const T* x = reinterpret_cast<const T*>(dy.x);
        const T* x2 = reinterpret_cast<const T*>(dy.x2);
        for (long long r = (long long)blockIdx.x * 4 + rl; r < M; r += (long long)gridDim.x * 4) {
            float v = Vec4<T>::ld1(x + r * N + c);
            if (dy.mode == B200SP_VT_DY) v = fmaf(p0, v, fmaf(p1, Vec4<T>::ld1(x2 + r * N + c), p2));
            a += v;
        }
    }
    s[rl][cl] = a;
    __syncthreads();
    if (threadIdx.x < 64 && c < N) atomicAdd(out + c, s[0][cl] + s[1][cl] + s[2][cl] + s[3][cl]);
}

__global__ void add_i64_kernel(int64_t* p, long long n, long long v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] += v;
}

// ---- RouterV2 (park2019.py:70-80) -------------------------------------------------------------
// out[b,h,w, (sy*2+sx)*Cr + c] = f_r(xr[b, 2h+sy, 2w+sx, c]);  out[b,h,w, 4Cr + c] = f_1(x1[b,h,w,c])
template <typename T>
__global__ void __launch_bounds__(EW_NT) reorg_cat_fwd_kernel(const b200sp_vtensor xr, const b200sp_vtensor x1, T* __restrict__ out,
                                                              int B, int h, int w, int Cr, int C1) {
    pdl_wait();
    pdl_trigger();
    const int Ct = 4 * Cr + C1, Ct4 = Ct / 4;
    const long long n4 = (long long)B * h * w * Ct4;
    for (long long i = (long long)blockIdx.x * EW_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * EW_NT) {
        const int c = (int)(i % Ct4) * 4;
        long long pix = i / Ct4;
        const int ww = (int)(pix % w); pix /= w;
        const int hh = (int)(pix % h);
        const int b = (int)(pix / h);
        float4 v;
        if (c < 4 * Cr) {
            const int q = c / Cr, cc = c % Cr, sy = q >> 1, sx = q & 1;
            v = vt_load4<T>(xr, ((size_t)(b * 2 * h + 2 * hh + sy) * (2 * w) + 2 * ww + sx) * Cr + cc, cc);
        } else {
            const int cc = c - 4 * Cr;
            v = vt_load4<T>(x1, ((size_t)(b * h + hh) * w + ww) * C1 + cc, cc);
        }
        Vec4<T>::st(out + i * 4, v);
    }
}
// backward: split dcat back and apply act'(z) of each producer BN (s1/s2 come from bn_bwd_reduce)
template <typename T>
__global__ void __launch_bounds__(EW_NT) reorg_cat_bwd_kernel(const T* __restrict__ dcat, T* __restrict__ g_r, T* __restrict__ g_1,
                                                              const b200sp_bnbwd bn_r, const b200sp_bnbwd bn_1,
                                                              int B, int h, int w, int Cr, int C1) {
    pdl_wait();
    pdl_trigger();
    const int Ct = 4 * Cr + C1, Ct4 = Ct / 4;
    const long long n4 = (long long)B * h * w * Ct4;
    for (long long i = (long long)blockIdx.x * EW_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * EW_NT) {
        const int c = (int)(i % Ct4) * 4;
        long long pix = i / Ct4;
        const int ww = (int)(pix % w); pix /= w;
        const int hh = (int)(pix % h);
        const int b = (int)(pix / h);
        const float4 d = Vec4<T>::ld(dcat + i * 4);
        const bool is_r = c < 4 * Cr;
        const b200sp_bnbwd& bn = is_r ? bn_r : bn_1;
        int cc; size_t off;
        if (is_r) {
            const int q = c / Cr, sy = q >> 1, sx = q & 1;
            cc = c % Cr;
            off = ((size_t)(b * 2 * h + 2 * hh + sy) * (2 * w) + 2 * ww + sx) * Cr + cc;
        } else {
            cc = c - 4 * Cr;
            off = ((size_t)(b * h + hh) * w + ww) * C1 + cc;
        }
        const float4 y = Vec4<T>::ld(reinterpret_cast<const T*>(bn.y) + off);
        const float4 sc = ldg4(bn.scale + cc), sh = ldg4(bn.shift + cc);
        float4 gq;
        gq.x = d.x * act_bwd(fmaf(y.x, sc.x, sh.x), bn.act); gq.y = d.y * act_bwd(fmaf(y.y, sc.y, sh.y), bn.act);
        gq.z = d.z * act_bwd(fmaf(y.z, sc.z, sh.z), bn.act); gq.w = d.w * act_bwd(fmaf(y.w, sc.w, sh.w), bn.act);
        Vec4<T>::st((is_r ? g_r : g_1) + off, gq);
    }
}

// ---- KRN head: logits[b,n] += sum_k a[b,k] * W[n,k],  k = (h*W + w)*C + c ---------------------
constexpr int HEAD_MAXN = 24;
// Second generation of the head forward (round 2; the first spent 72 us on 5-step shuffle reductions per (image, output)).
// One CTA = a 128-wide slice of the 50176-long reduction: the activated inputs [B][128] and the weights [N][128] of the slice
// are staged once (16-byte loads), then thread o owns logits (b, n) = (o / N, o % N), o += 256, as a 128-long dot product
// read from shared memory in 16-byte pieces (row pitch 132 floats: 8 consecutive rows cover all banks), one atomicAdd each.
constexpr int HF_KC = 128, HF_LD = HF_KC + 4, HF_MAXB = 64;
template <typename T>
__global__ void __launch_bounds__(EW_NT) head_fwd2_kernel(const b200sp_vtensor x, const float* __restrict__ w, float* __restrict__ logits,
                                                          int B, int HWC, int C, int N) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float hf_smem[];
    float* s_x = hf_smem;                       // [B][HF_LD]
    float* s_w = hf_smem + (size_t)B * HF_LD;   // [N][HF_LD]
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * HF_KC;
    const ActP ap = act_params(x.act);
    for (int i = tid; i < B * (HF_KC / 4); i += EW_NT) {
        const int b = i / (HF_KC / 4), k4 = (i - b * (HF_KC / 4)) * 4, k = k0 + k4;
        float4 v = f4zero();
        if (k < HWC) {
            v = Vec4<T>::ld(reinterpret_cast<const T*>(x.x) + (size_t)b * HWC + k);
            if (x.mode == B200SP_VT_BNACT) {
                const int c = k % C;
                const float4 sc = ldg4(x.p0 + c), sh = ldg4(x.p1 + c);
                v = make_float4(act_fwd(fmaf(v.x, sc.x, sh.x), ap), act_fwd(fmaf(v.y, sc.y, sh.y), ap),
                                act_fwd(fmaf(v.z, sc.z, sh.z), ap), act_fwd(fmaf(v.w, sc.w, sh.w), ap));
            }
        }
        *reinterpret_cast<float4*>(s_x + b * HF_LD + k4) = v;
    }
    for (int i = tid; i < N * (HF_KC / 4); i += EW_NT) {
        const int n = i / (HF_KC / 4), k4 = (i - n * (HF_KC / 4)) * 4, k = k0 + k4;
        float4 v = f4zero();
        if (k < HWC) v = ldg4(w + (size_t)n * HWC + k);
        *reinterpret_cast<float4*>(s_w + n * HF_LD + k4) = v;
    }
    __syncthreads();
    for (int o = tid; o < B * N; o += EW_NT) {
        const int b = o / N, n = o - b * N;
        const float4* xr = reinterpret_cast<const float4*>(s_x + b * HF_LD);
        const float4* wr = reinterpret_cast<const float4*>(s_w + n * HF_LD);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int j = 0; j < HF_KC / 4; ++j) {
            const float4 xv = xr[j], wv = wr[j];
            a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1);
            a0 = fmaf(xv.z, wv.z, a0); a1 = fmaf(xv.w, wv.w, a1);
        }
        atomicAdd(logits + o, a0 + a1);
    }
}

constexpr int HEAD_BT = 8;       // batch rows per smem pass
template <typename T>
__global__ void __launch_bounds__(EW_NT) head_fwd_kernel(const b200sp_vtensor x, const float* __restrict__ w, float* __restrict__ logits,
                                                         int B, int HWC, int C, int N) {
    __shared__ float s_part[8][HEAD_BT][HEAD_MAXN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = blockIdx.x * EW_NT + tid;
    const bool ok = k < HWC;
    float wv[HEAD_MAXN];
#pragma unroll
    for (int n = 0; n < HEAD_MAXN; ++n) wv[n] = (ok && n < N) ? __ldg(w + (size_t)n * HWC + k) : 0.f;
    float sc = 1.f, sh = 0.f;
    const int c = ok ? k % C : 0;
    if (x.mode == B200SP_VT_BNACT) { sc = x.p0[c]; sh = x.p1[c]; }
    for (int b0 = 0; b0 < B; b0 += HEAD_BT) {
#pragma unroll
        for (int bb = 0; bb < HEAD_BT; ++bb) {
            const int b = b0 + bb;
            float a = 0.f;
            if (ok && b < B) {
                a = Vec4<T>::ld1(reinterpret_cast<const T*>(x.x) + (size_t)b * HWC + k);
                if (x.mode == B200SP_VT_BNACT) a = act_fwd(fmaf(a, sc, sh), x.act);
            }
#pragma unroll
            for (int n = 0; n < HEAD_MAXN; ++n) {
                float p = a * wv[n];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) s_part[warp][bb][n] = p;
            }
        }
        __syncthreads();
        if (tid < HEAD_BT * HEAD_MAXN) {
            const int bb = tid / HEAD_MAXN, n = tid % HEAD_MAXN, b = b0 + bb;
            if (b < B && n < N) {
                float s = 0.f;
#pragma unroll
                for (int wp = 0; wp < 8; ++wp) s += s_part[wp][bb][n];
                atomicAdd(logits + b * N + n, s);
            }
        }
        __syncthreads();
    }
}

__global__ void head_bias_kernel(const float* __restrict__ bias, float* __restrict__ logits, int B, int N) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * N) logits[i] = bias[i % N];
}

// park2019.py:142-160.  One CTA.  target [B,2,nk]; logits col 2i = x_i, col 2i+1 = y_i.
__global__ void krn_loss_kernel(const float* __restrict__ logits, const float* __restrict__ target, float* __restrict__ loss3,
                                float* __restrict__ dlogits, float* __restrict__ dbias, const float* __restrict__ loss_scale,
                                int B, int N) {
    pdl_wait();
    pdl_trigger();
    __shared__ float s_lx[EW_NT], s_ly[EW_NT];
    const int tid = threadIdx.x, nk = N / 2;
    const float ls = loss_scale ? loss_scale[0] : 1.f;
    float lx = 0.f, ly = 0.f;
    for (int i = tid; i < B * N; i += EW_NT) {
        const int b = i / N, n = i % N, kp = n >> 1, isy = n & 1;
        const float d = logits[i] - target[(b * 2 + isy) * nk + kp];
        if (isy) ly = fmaf(d, d, ly); else lx = fmaf(d, d, lx);
        dlogits[i] = 2.f * d / (float)B * ls;
    }
    s_lx[tid] = lx; s_ly[tid] = ly;
    __syncthreads();
    for (int o = EW_NT / 2; o > 0; o >>= 1) {
        if (tid < o) { s_lx[tid] += s_lx[tid + o]; s_ly[tid] += s_ly[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const float a = s_lx[0] / (float)B, b = s_ly[0] / (float)B;
        loss3[0] = a + b; loss3[1] = a; loss3[2] = b;
    }
    if (dbias && tid < N) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dlogits[b * N + tid];     // written by this CTA above
        dbias[tid] += s;
    }
}

// head backward: thread per k.  g[b,k] = (sum_n dl[b,n] W[n,k]) * act'(z);  dW[n,k] += sum_b dl[b,n] a[b,k]
template <typename T>
__global__ void __launch_bounds__(EW_NT) head_bwd_kernel(const float* __restrict__ dlogits, const b200sp_vtensor x,
                                                         const float* __restrict__ w, T* __restrict__ g, float* __restrict__ dw,
                                                         float* __restrict__ dbias, const b200sp_bnbwd bn, const int has_bn,
                                                         int B, int HWC, int C, int N) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float s_dl[];     // [B][HEAD_MAXN]
    const int tid = threadIdx.x;
    for (int i = tid; i < B * HEAD_MAXN; i += EW_NT) {
        const int b = i / HEAD_MAXN, n = i % HEAD_MAXN;
        s_dl[i] = n < N ? dlogits[b * N + n] : 0.f;
    }
    __syncthreads();
    if (dbias && blockIdx.x == 0 && tid < N) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += s_dl[b * HEAD_MAXN + tid];
        dbias[tid] += s;
    }
    const int k = blockIdx.x * EW_NT + tid;
    if (k >= HWC) return;
    const int c = k % C;
    float wv[HEAD_MAXN], dwv[HEAD_MAXN];
#pragma unroll
    for (int n = 0; n < HEAD_MAXN; ++n) { wv[n] = n < N ? __ldg(w + (size_t)n * HWC + k) : 0.f; dwv[n] = 0.f; }
    float sc = 1.f, sh = 0.f, mu = 0.f, rs = 0.f;
    const bool bnact = x.mode == B200SP_VT_BNACT;
    if (bnact) { sc = x.p0[c]; sh = x.p1[c]; }
    const bool stats = has_bn && bn.s1 != nullptr;
    if (stats) { mu = bn.mean[c]; rs = bn.rstd[c]; }
    float s1 = 0.f, s2 = 0.f;
    for (int b = 0; b < B; ++b) {
        const float y = Vec4<T>::ld1(reinterpret_cast<const T*>(x.x) + (size_t)b * HWC + k);
        const float z = bnact ? fmaf(y, sc, sh) : y;
        const float a = bnact ? act_fwd(z, x.act) : y;
        float da = 0.f;
#pragma unroll
        for (int n = 0; n < HEAD_MAXN; ++n) {
            const float d = s_dl[b * HEAD_MAXN + n];
            da = fmaf(d, wv[n], da);
            dwv[n] = fmaf(d, a, dwv[n]);
        }
        const float gv = bnact ? da * act_bwd(z, x.act) : da;
        Vec4<T>::st1(g + (size_t)b * HWC + k, gv);
        s1 += gv;
        s2 = fmaf(gv, (y - mu) * rs, s2);
    }
#pragma unroll
    for (int n = 0; n < HEAD_MAXN; ++n) if (n < N) dw[(size_t)n * HWC + k] += dwv[n];
    if (stats) { atomicAdd(bn.s1 + c, (double)s1); atomicAdd(bn.s2 + c, (double)s2); }
}

inline int ew_grid(long long n) {
    long long g = (n + EW_NT - 1) / EW_NT;
    long long cap = (long long)NUM_SMS * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

extern "C" int b200sp_bn_apply(const void* y, const float* scale, const float* shift, const void* residual,
                               int act, void* out, int64_t M, int C, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    if (C % 4) return B200SP_EINVAL;
    const long long n4 = M * (C / 4);
    if (dtype == B200SP_F32)
        b200sp_launch_pdl(bn_apply_kernel<float>, dim3(ew_grid((n4 + 3) / 4)), dim3(EW_NT), 0, (cudaStream_t)stream, (const float*)y, scale, shift,
                          (const float*)residual, act, (float*)out, n4, C / 4);
    else
        bn_apply_kernel<bf16><<<ew_grid(n4), EW_NT, 0, (cudaStream_t)stream>>>((const bf16*)y, scale, shift, (const bf16*)residual,
                                                                             act, (bf16*)out, n4, C / 4);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_bn_eval_affine(const float* gamma, const float* beta, const float* rmean, const float* rvar,
                                     float eps, float* scale, float* shift, int64_t n, void* stream) {
    bn_eval_affine_kernel<<<ew_grid(n), EW_NT, 0, (cudaStream_t)stream>>>(gamma, beta, rmean, rvar, eps, scale, shift, n);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_bn_fwd_finalize(const b200sp_bnfwd* bn, int C, double count, void* stream) {
    bn_fwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(*bn, C, count);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
extern "C" int b200sp_bn_bwd_finalize(const b200sp_bnbwd* bn, int C, double count, void* stream) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(*bn, C, count);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_bn_bwd_reduce(const void* g, const b200sp_bnbwd* bn, int64_t M, int C, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    const int gy = ceil_div(C, 64);
    long long gx = (M + 3) / 4;
    const long long cap = (NUM_SMS * 8 + gy - 1) / gy;
    if (gx > cap) gx = cap;
    if (dtype == B200SP_F32) bn_bwd_reduce_kernel<float><<<dim3((unsigned)gx, gy), EW_NT, 0, (cudaStream_t)stream>>>((const float*)g, *bn, M, C);
    else bn_bwd_reduce_kernel<bf16><<<dim3((unsigned)gx, gy), EW_NT, 0, (cudaStream_t)stream>>>((const bf16*)g, *bn, M, C);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_colsum_f32(const b200sp_vtensor* dy, float* out, int M, int N, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    const int gy = ceil_div(N, 64);
    long long gx = ((long long)M + 3) / 4;
    const long long cap = (NUM_SMS * 4 + gy - 1) / gy;
    if (gx > cap) gx = cap;
    if (dtype == B200SP_F32) colsum_kernel<float><<<dim3((unsigned)gx, gy), EW_NT, 0, (cudaStream_t)stream>>>(*dy, out, M, N);
    else colsum_kernel<bf16><<<dim3((unsigned)gx, gy), EW_NT, 0, (cudaStream_t)stream>>>(*dy, out, M, N);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_add_i64(int64_t* p, int64_t n, int64_t v, void* stream) {
    add_i64_kernel<<<ew_grid(n), EW_NT, 0, (cudaStream_t)stream>>>(p, n, v);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_reorg_cat_fwd(const b200sp_vtensor* xr, const b200sp_vtensor* x1, void* out,
                                    int B, int h, int w, int Cr, int C1, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    if (Cr % 4 || C1 % 4) return B200SP_EINVAL;
    const long long n4 = (long long)B * h * w * ((4 * Cr + C1) / 4);
    if (dtype == B200SP_F32) b200sp_launch_pdl(reorg_cat_fwd_kernel<float>, dim3(ew_grid(n4)), dim3(EW_NT), 0, (cudaStream_t)stream, *xr, *x1, (float*)out, B, h, w, Cr, C1);
    else reorg_cat_fwd_kernel<bf16><<<ew_grid(n4), EW_NT, 0, (cudaStream_t)stream>>>(*xr, *x1, (bf16*)out, B, h, w, Cr, C1);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_reorg_cat_bwd(const void* dcat, void* g_r, void* g_1, const b200sp_bnbwd* bn_r,
                                    const b200sp_bnbwd* bn_1, int B, int h, int w, int Cr, int C1, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    if (Cr % 4 || C1 % 4) return B200SP_EINVAL;
    const long long n4 = (long long)B * h * w * ((4 * Cr + C1) / 4);
    if (dtype == B200SP_F32)
        b200sp_launch_pdl(reorg_cat_bwd_kernel<float>, dim3(ew_grid(n4)), dim3(EW_NT), 0, (cudaStream_t)stream, (const float*)dcat, (float*)g_r, (float*)g_1,
                          *bn_r, *bn_1, B, h, w, Cr, C1);
    else
        reorg_cat_bwd_kernel<bf16><<<ew_grid(n4), EW_NT, 0, (cudaStream_t)stream>>>((const bf16*)dcat, (bf16*)g_r, (bf16*)g_1,
                                                                                  *bn_r, *bn_1, B, h, w, Cr, C1);
    B200SP_COUNT_LAUNCH();
    if (int rc = b200sp_bn_bwd_reduce(g_r, bn_r, (int64_t)B * 4 * h * w, Cr, dtype, stream)) return rc;
    return b200sp_bn_bwd_reduce(g_1, bn_1, (int64_t)B * h * w, C1, dtype, stream);
}

extern "C" int b200sp_head_bias(const float* bias, float* logits, int B, int N, void* stream) {
    b200sp_launch_pdl(head_bias_kernel, dim3(ceil_div(B * N, 256)), dim3(256), 0, (cudaStream_t)stream, bias, logits, B, N);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_head_fwd(const b200sp_vtensor* x, const float* w, float* logits,
                               int B, int HWC, int C, int N, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    if (N > HEAD_MAXN || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    if (dtype == B200SP_F32 && B <= HF_MAXB && HWC % 4 == 0 && C % 4 == 0 && x->act != B200SP_ACT_SIGMOID &&
        (((uintptr_t)x->x | (uintptr_t)w) & 15) == 0) {
        const size_t smem = sizeof(float) * (size_t)(B + N) * HF_LD;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(head_fwd2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (HF_MAXB + HEAD_MAXN) * HF_LD));
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        b200sp_launch_pdl(head_fwd2_kernel<float>, dim3(ceil_div(HWC, HF_KC)), dim3(EW_NT), smem, (cudaStream_t)stream, *x, w, logits, B, HWC, C, N);
        B200SP_COUNT_LAUNCH();
        B200SP_RETURN_LAST();
    }
    if (dtype == B200SP_F32) head_fwd_kernel<float><<<ceil_div(HWC, EW_NT), EW_NT, 0, (cudaStream_t)stream>>>(*x, w, logits, B, HWC, C, N);
    else head_fwd_kernel<bf16><<<ceil_div(HWC, EW_NT), EW_NT, 0, (cudaStream_t)stream>>>(*x, w, logits, B, HWC, C, N);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_krn_loss(const float* logits, const float* target, float* loss3, float* dlogits,
                               float* dbias, const float* loss_scale, int B, int N, void* stream) {
    if (N > EW_NT || N % 2) return B200SP_EINVAL;
    b200sp_launch_pdl(krn_loss_kernel, dim3(1), dim3(EW_NT), 0, (cudaStream_t)stream, logits, target, loss3, dlogits, dbias, loss_scale, B, N);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_head_bwd(const float* dlogits, const b200sp_vtensor* x, const float* w, void* g, float* dw,
                               float* dbias, const b200sp_bnbwd* bn, int B, int HWC, int C, int N, int dtype, void* stream) {
    if (dtype != B200SP_F32 && dtype != B200SP_BF16) return B200SP_ENOSYS;
    if (N > HEAD_MAXN || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    b200sp_bnbwd b = {};
    if (bn) b = *bn;
    const size_t smem = (size_t)B * HEAD_MAXN * sizeof(float);
    if (smem > 48 * 1024) return B200SP_EINVAL;
    if (dtype == B200SP_F32)
        b200sp_launch_pdl(head_bwd_kernel<float>, dim3(ceil_div(HWC, EW_NT)), dim3(EW_NT), smem, (cudaStream_t)stream, dlogits, *x, w, (float*)g, dw, dbias, b,
                          (int)(bn != nullptr), B, HWC, C, N);
    else
        head_bwd_kernel<bf16><<<ceil_div(HWC, EW_NT), EW_NT, smem, (cudaStream_t)stream>>>(dlogits, *x, w, (bf16*)g, dw, dbias, b, bn != nullptr,
                                                                                           B, HWC, C, N);
    B200SP_COUNT_LAUNCH();
    if (bn && bn->s1) return b200sp_bn_bwd_finalize(bn, C, (double)B * (HWC / C), stream);
    B200SP_RETURN_LAST();
}
