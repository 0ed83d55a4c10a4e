// spn.cu -- the non-GEMM kernels of the SPN (AlexNet) path, src/nets/spn.py:37-143:
// im2col / col2im around the tcgen05 GEMMs (dense convolutions 11x11 s4, 5x5 p2 g2, 3x3 p1 [g2] are run as
// GEMMs over an explicit patch matrix; the patch matrix is fp32 so the 3xTF32 GEMM keeps fp32-grade accuracy,
// which the "attitude-class argmax bit-exact" bar needs), MaxPool2d(3,2) + LocalResponseNorm(2, 2e-5, .75, 1)
// forward/backward, dropout, and the TF-style soft-target cross entropy (spn.py:37-48).  All NHWC fp32.
#include "common.cuh"

namespace {

constexpr int SP_NT = 256;

inline int sp_grid(long long n) {
    long long g = (n + SP_NT - 1) / SP_NT;
    const long long cap = (long long)NUM_SMS * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

// col[m][(kh*k + kw)*Cg + c] = x[b, oh*s - p + kh, ow*s - p + kw, c_off + c]  (0 outside);  K padded to Kp with zeros.
// One thread = one 16-byte store (4 consecutive columns); when Cg % 4 == 0 (every layer but conv1) the 4 columns are
// 4 consecutive channels of one tap, i.e. one 16-byte NHWC load.
__global__ void __launch_bounds__(SP_NT) im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W, int C,
                                                      int c_off, int Cg, int k, int s, int p, int Ho, int Wo, int Kp, int nchw) {
    const int K4 = Kp >> 2, Kreal = k * k * Cg;
    const long long n = (long long)B * Ho * Wo * K4;
    const bool vec = !nchw && (Cg & 3) == 0 && (C & 3) == 0 && (c_off & 3) == 0;
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const int kk = (int)(i % K4) * 4;
        const long long m = i / K4;
        const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
        float4 v = f4zero();
        if (vec) {
            if (kk < Kreal) {
                const int c = kk % Cg, t = kk / Cg, kw = t % k, kh = t / k;
                const int ih = oh * s - p + kh, iw = ow * s - p + kw;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = ldg4(x + (((size_t)b * H + ih) * W + iw) * C + c_off + c);
            }
        } else {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = kk + j;
                e[j] = 0.f;
                if (q < Kreal) {
                    const int c = q % Cg, t = q / Cg, kw = t % k, kh = t / k;
                    const int ih = oh * s - p + kh, iw = ow * s - p + kw;
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
                        e[j] = nchw ? __ldg(x + (((size_t)b * C + c_off + c) * H + ih) * W + iw)
                                    : __ldg(x + (((size_t)b * H + ih) * W + iw) * C + c_off + c);
                }
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4*>(col + i * 4) = v;
    }
}

// dx[b,ih,iw,c_off+c] = mask * sum_{kh,kw} dcol[m(oh,ow)][(kh*k+kw)*Cg + c],  oh = (ih+p-kh)/s  (gather: no atomics)
// mask = (a[b,ih,iw,c_off+c] > 0) when `a` (the post-ReLU activation that produced x) is given
__global__ void __launch_bounds__(SP_NT) col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, const float* __restrict__ a,
                                                      int B, int H, int W, int C, int c_off, int Cg, int k, int s, int p, int Ho, int Wo,
                                                      int Kp) {
    const long long n = (long long)B * H * W * Cg;
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const int c = (int)(i % Cg);
        long long r = i / Cg;
        const int iw = (int)(r % W); r /= W;
        const int ih = (int)(r % H);
        const int b = (int)(r / H);
        float acc = 0.f;
        for (int kh = 0; kh < k; ++kh) {
            const int th = ih + p - kh;
            if (th < 0 || th % s) continue;
            const int oh = th / s;
            if (oh >= Ho) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int tw = iw + p - kw;
                if (tw < 0 || tw % s) continue;
                const int ow = tw / s;
                if (ow >= Wo) continue;
                acc += __ldg(dcol + (((size_t)b * Ho + oh) * Wo + ow) * Kp + (kh * k + kw) * Cg + c);
            }
        }
        const size_t o = (((size_t)b * H + ih) * W + iw) * C + c_off + c;
        if (a && !(a[o] > 0.f)) acc = 0.f;
        dx[o] = acc;
    }
}

// MaxPool2d(3, 2) then optional LocalResponseNorm(size 2): out = p / (1 + alpha/2 (p[c-1]^2 + p[c]^2))^beta
__global__ void __launch_bounds__(SP_NT) pool_lrn_fwd_kernel(const float* __restrict__ x, float* __restrict__ pooled, float* __restrict__ out,
                                                            uint8_t* __restrict__ amax, int B, int H, int W, int C, int Ho, int Wo, int lrn,
                                                            float alpha, float beta) {
    const long long n = (long long)B * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int b = (int)(r / Ho);
        int am = 0;
        auto pool = [&](int cc) {
            float m = -3.4e38f;
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = __ldg(x + (((size_t)b * H + oh * 2 + kh) * W + ow * 2 + kw) * C + cc);
                    if (v > m) { m = v; if (cc == c) am = kh * 3 + kw; }      // first maximum wins (torch max_pool2d)
                }
            return m;
        };
        const float pc = pool(c);
        if (pooled) pooled[i] = pc;
        if (amax) amax[i] = (uint8_t)am;
        if (lrn) {
            const float pm = c > 0 ? pool(c - 1) : 0.f;
            const float d = 1.f + 0.5f * alpha * (pm * pm + pc * pc);
            out[i] = pc * powf(d, -beta);
        } else {
            out[i] = pc;
        }
    }
}

// gradient wrt the pooled tensor of the LRN:  dp_c = g_c d_c^-b - alpha*b*p_c (g_c p_c d_c^(-b-1) + g_{c+1} p_{c+1} d_{c+1}^(-b-1))
__global__ void __launch_bounds__(SP_NT) lrn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ pooled, float* __restrict__ dp,
                                                       long long npix, int C, float alpha, float beta) {
    const long long n = npix * C;
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const int c = (int)(i % C);
        const float pc = pooled[i];
        const float pm = c > 0 ? pooled[i - 1] : 0.f;
        const float dc = 1.f + 0.5f * alpha * (pm * pm + pc * pc);
        float t = g[i] * pc * powf(dc, -beta - 1.f);
        float acc = g[i] * powf(dc, -beta);
        if (c + 1 < C) {
            const float pn = pooled[i + 1];
            const float dn = 1.f + 0.5f * alpha * (pc * pc + pn * pn);
            t += g[i + 1] * pn * powf(dn, -beta - 1.f);
        }
        dp[i] = acc - alpha * beta * pc * t;
    }
}

// MaxPool2d(3,2) backward in gather form: dx = relu'(x) * sum over the (<= 4) windows that contain this element and
// whose recorded first-maximum index (amax, written by the forward) is this element.
__global__ void __launch_bounds__(SP_NT) pool_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ x, const uint8_t* __restrict__ amax,
                                                        float* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo, int relu_mask) {
    const long long n = (long long)B * H * W * C;
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int iw = (int)(r % W); r /= W;
        const int ih = (int)(r % H);
        const int b = (int)(r / H);
        float acc = 0.f;
        if (!(relu_mask && !(x[i] > 0.f))) {
            for (int oh = max(0, (ih - 1) / 2); oh <= min(Ho - 1, ih / 2); ++oh) {
                const int kh = ih - oh * 2;
                if (kh < 0 || kh > 2) continue;
                for (int ow = max(0, (iw - 1) / 2); ow <= min(Wo - 1, iw / 2); ++ow) {
                    const int kw = iw - ow * 2;
                    if (kw < 0 || kw > 2) continue;
                    const size_t o = (((size_t)b * Ho + oh) * Wo + ow) * C + c;
                    if (amax[o] == kh * 3 + kw) acc += __ldg(dp + o);
                }
            }
        }
        dx[i] = acc;
    }
}

// out = x * mask / (1-p), mask from a counter-based hash of (seed, element index); mask (uint8) saved for backward
__device__ __forceinline__ uint32_t mix32(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
    return a;
}
__global__ void __launch_bounds__(SP_NT) dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ mask,
                                                           long long n, float p, uint32_t seed_lo, uint32_t seed_hi,
                                                           const unsigned long long* __restrict__ ctr) {
    const float scale = 1.f / (1.f - p);
    if (ctr) {            // device-resident step counter: a captured CUDA graph draws a fresh mask on every replay
        const unsigned long long c = *ctr * 0x9E3779B97F4A7C15ull;
        seed_lo ^= mix32((uint32_t)c);
        seed_hi ^= mix32((uint32_t)(c >> 32) + 0x85ebca6bu);
    }
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT) {
        const uint32_t h = mix32((uint32_t)i ^ mix32(seed_lo + 0x9e3779b9u * (uint32_t)(i >> 32)) ^ seed_hi);
        const bool keep = (h >> 8) * (1.f / 16777216.f) >= p;
        mask[i] = keep;
        out[i] = keep ? x[i] * scale : 0.f;
    }
}
__global__ void __launch_bounds__(SP_NT) dropout_bwd_kernel(float* __restrict__ g, const uint8_t* __restrict__ mask, long long n, float scale) {
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT)
        g[i] = mask[i] ? g[i] * scale : 0.f;
}

// one CTA per row:  loss[b] = -sum_j t_j (z_j - lse);  dz_j = weight/B * (softmax_j * sum(t) - t_j)
__global__ void __launch_bounds__(SP_NT) soft_ce_kernel(const float* __restrict__ z, const float* __restrict__ t, float* __restrict__ loss_rows,
                                                       float* __restrict__ dz, int N, float weight_over_B) {
    __shared__ float s[SP_NT];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* zr = z + (size_t)b * N;
    const float* tr = t + (size_t)b * N;
    float m = -3.4e38f;
    for (int j = tid; j < N; j += SP_NT) m = fmaxf(m, zr[j]);
    s[tid] = m; __syncthreads();
    for (int o = SP_NT / 2; o > 0; o >>= 1) { if (tid < o) s[tid] = fmaxf(s[tid], s[tid + o]); __syncthreads(); }
    m = s[0]; __syncthreads();
    float se = 0.f, st = 0.f, stz = 0.f;
    for (int j = tid; j < N; j += SP_NT) { se += expf(zr[j] - m); st += tr[j]; stz += tr[j] * zr[j]; }
    s[tid] = se; __syncthreads();
    for (int o = SP_NT / 2; o > 0; o >>= 1) { if (tid < o) s[tid] += s[tid + o]; __syncthreads(); }
    se = s[0]; __syncthreads();
    s[tid] = st; __syncthreads();
    for (int o = SP_NT / 2; o > 0; o >>= 1) { if (tid < o) s[tid] += s[tid + o]; __syncthreads(); }
    st = s[0]; __syncthreads();
    s[tid] = stz; __syncthreads();
    for (int o = SP_NT / 2; o > 0; o >>= 1) { if (tid < o) s[tid] += s[tid + o]; __syncthreads(); }
    stz = s[0];
    const float lse = m + logf(se);
    if (tid == 0) loss_rows[b] = st * lse - stz;
    if (dz) {
        const float inv = 1.f / se;
        for (int j = tid; j < N; j += SP_NT) dz[(size_t)b * N + j] = weight_over_B * (expf(zr[j] - m) * inv * st - tr[j]);
    }
}
// loss2[0] = mean(rows_c), loss2[1] = mean(rows_r)
__global__ void soft_ce_mean_kernel(const float* __restrict__ rows_c, const float* __restrict__ rows_r, float* __restrict__ loss2, int B) {
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < B; ++i) { a += rows_c[i]; b += rows_r[i]; }
        loss2[0] = a / (float)B;
        loss2[1] = b / (float)B;
    }
}

// y = act(y + bias[n]) in place (the epilogue of a split-K FC forward)
__global__ void __launch_bounds__(SP_NT) bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n4, int N4, int relu) {
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * SP_NT) {
        float4 v = *reinterpret_cast<float4*>(y + i * 4);
        const float4 b = ldg4(bias + (i % N4) * 4);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(y + i * 4) = v;
    }
}

__global__ void __launch_bounds__(SP_NT) relu_mask_kernel(float* __restrict__ g, const float* __restrict__ a, long long n) {
    for (long long i = (long long)blockIdx.x * SP_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SP_NT)
        if (!(a[i] > 0.f)) g[i] = 0.f;
}

}  // namespace

extern "C" int b200sp_im2col(const float* x, float* col, int B, int H, int W, int C, int c_off, int Cg, int k, int stride, int pad,
                             int Kp, int nchw, void* stream) {
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    if (Kp < k * k * Cg || (Kp & 3) || Ho < 1 || Wo < 1) return B200SP_EINVAL;
    im2col_kernel<<<sp_grid((long long)B * Ho * Wo * (Kp / 4)), SP_NT, 0, (cudaStream_t)stream>>>(x, col, B, H, W, C, c_off, Cg, k, stride, pad, Ho, Wo, Kp, nchw);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_col2im(const float* dcol, float* dx, const float* act_mask, int B, int H, int W, int C, int c_off, int Cg, int k,
                             int stride, int pad, int Kp, void* stream) {
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    col2im_kernel<<<sp_grid((long long)B * H * W * Cg), SP_NT, 0, (cudaStream_t)stream>>>(dcol, dx, act_mask, B, H, W, C, c_off, Cg, k, stride, pad, Ho, Wo, Kp);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_pool_lrn_fwd(const float* x, float* pooled, float* out, uint8_t* amax, int B, int H, int W, int C, int lrn, float alpha,
                                   float beta, void* stream) {
    const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1;
    pool_lrn_fwd_kernel<<<sp_grid((long long)B * Ho * Wo * C), SP_NT, 0, (cudaStream_t)stream>>>(x, pooled, out, amax, B, H, W, C, Ho, Wo, lrn, alpha, beta);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_pool_lrn_bwd(const float* g_out, const float* pooled, const float* x, const uint8_t* amax, float* scratch, float* dx,
                                   int B, int H, int W, int C, int lrn, float alpha, float beta, int relu_mask, void* stream) {
    if (!amax) return B200SP_EINVAL;
    const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1;
    const float* dp = g_out;
    if (lrn) {
        lrn_bwd_kernel<<<sp_grid((long long)B * Ho * Wo * C), SP_NT, 0, (cudaStream_t)stream>>>(g_out, pooled, scratch, (long long)B * Ho * Wo, C, alpha, beta);
        B200SP_COUNT_LAUNCH();
        dp = scratch;
    }
    pool_bwd_kernel<<<sp_grid((long long)B * H * W * C), SP_NT, 0, (cudaStream_t)stream>>>(dp, x, amax, dx, B, H, W, C, Ho, Wo, relu_mask);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_dropout_fwd(const float* x, float* out, uint8_t* mask, int64_t n, float p, uint64_t seed, void* stream) {
    if (p < 0.f || p >= 1.f) return B200SP_EINVAL;
    dropout_fwd_kernel<<<sp_grid(n), SP_NT, 0, (cudaStream_t)stream>>>(x, out, mask, n, p, (uint32_t)seed, (uint32_t)(seed >> 32), nullptr);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_dropout_fwd_ctr(const float* x, float* out, uint8_t* mask, int64_t n, float p, uint64_t seed, const int64_t* counter,
                                      void* stream) {
    if (p < 0.f || p >= 1.f || !counter) return B200SP_EINVAL;
    dropout_fwd_kernel<<<sp_grid(n), SP_NT, 0, (cudaStream_t)stream>>>(x, out, mask, n, p, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                      reinterpret_cast<const unsigned long long*>(counter));
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_dropout_bwd(float* g, const uint8_t* mask, int64_t n, float p, void* stream) {
    dropout_bwd_kernel<<<sp_grid(n), SP_NT, 0, (cudaStream_t)stream>>>(g, mask, n, 1.f / (1.f - p));
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_soft_ce(const float* logits, const float* target, float* loss_rows, float* dlogits, int B, int N, float weight,
                              void* stream) {
    soft_ce_kernel<<<B, SP_NT, 0, (cudaStream_t)stream>>>(logits, target, loss_rows, dlogits, N, weight / (float)B);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_soft_ce_mean(const float* rows_c, const float* rows_r, float* loss2, int B, void* stream) {
    soft_ce_mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rows_c, rows_r, loss2, B);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_relu_mask(float* g, const float* a, int64_t n, void* stream) {
    relu_mask_kernel<<<sp_grid(n), SP_NT, 0, (cudaStream_t)stream>>>(g, a, n);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_bias_act(float* y, const float* bias, int M, int N, int relu, void* stream) {
    if (N % 4) return B200SP_EINVAL;
    const long long n4 = (long long)M * (N / 4);
    bias_act_kernel<<<sp_grid(n4), SP_NT, 0, (cudaStream_t)stream>>>(y, bias, n4, N / 4, relu);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
