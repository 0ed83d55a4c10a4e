// styleaug.cu -- the memory-bound half of the Ghiasi style-transfer forward (src/styleaug/ghiasi.py:6-135,
// src/styleaug/styleAugmentor.py:44-68): InstanceNorm finalisation, the fused
// normalise + conditional affine + ReLU (+ residual) pass that WRITES THE NEXT CONVOLUTION'S INPUT in the
// layout convtc.cu consumes (reflection padding, nearest x2 upsampling and the stride-2 phase split are
// index arithmetic on the destination side, so they cost no extra pass), and the style-embedding algebra.
#include "common.cuh"

namespace {

constexpr int SA_NT = 256;

__device__ __forceinline__ int reflect(int u, int n) {       // ReflectionPad2d index (pad < n)
    u = u < 0 ? -u : u;
    return u >= n ? 2 * (n - 1) - u : u;
}

__global__ void __launch_bounds__(SA_NT) sa_prep_kernel(const float* __restrict__ x, bf16* __restrict__ plane, int B, int H, int W,
                                                        int pad, int Cd) {
    const int Hd = H + 2 * pad, Wd = W + 2 * pad;
    const long long n = (long long)B * Hd * Wd;
    for (long long i = (long long)blockIdx.x * SA_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SA_NT) {
        const int xx = (int)(i % Wd), yy = (int)((i / Wd) % Hd), b = (int)(i / ((long long)Wd * Hd));
        const int sy = reflect(yy - pad, H), sx = reflect(xx - pad, W);
        const float* src = x + ((size_t)b * 3 * H + sy) * W + sx;
        bf16* dst = plane + (size_t)i * Cd;
        const size_t cs = (size_t)H * W;
        for (int c = 0; c < Cd; ++c) dst[c] = __float2bfloat16_rn(c < 3 ? __ldg(src + c * cs) : 0.f);
    }
}

__global__ void __launch_bounds__(SA_NT) in_finalize_kernel(float* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int gb_stride, float* __restrict__ scale,
                                                            float* __restrict__ shift, int B, int C, int Np, float inv_hw, float eps) {
    const int i = blockIdx.x * SA_NT + threadIdx.x;
    if (i >= B * Np) return;
    const int b = i / Np, c = i - b * Np;
    float* s = stats + (size_t)b * 2 * Np;
    const float sum = s[c], sq = s[Np + c];
    s[c] = 0.f;
    s[Np + c] = 0.f;
    if (c >= C) return;
    const float mean = sum * inv_hw;
    const float var = fmaxf(sq * inv_hw - mean * mean, 0.f);          // biased variance (instance_norm)
    const float rstd = rsqrtf(var + eps);
    const float ga = gamma ? gamma[(size_t)b * gb_stride + c] : 1.f;
    const float be = beta ? beta[(size_t)b * gb_stride + c] : 0.f;
    const float sc = ga * rstd;
    scale[b * C + c] = sc;
    shift[b * C + c] = be - mean * sc;
}

// one thread = 8 channels of one destination pixel (one 16-byte bf16 store)
__global__ void __launch_bounds__(SA_NT) in_apply_kernel(const b200sp_in_apply_desc d) {
    const int G = d.Cd / 8;
    const long long per_plane = (long long)d.B * d.Hd * d.Wd * G;
    const long long n = per_plane * d.ps * d.ps;
    const int Hup = d.Hs * d.up, Wup = d.Ws * d.up;
    for (long long i = (long long)blockIdx.x * SA_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SA_NT) {
        const int q = (int)(i / per_plane);
        long long r = i - q * per_plane;
        const int cg = (int)(r % G); r /= G;
        const int x = (int)(r % d.Wd); r /= d.Wd;
        const int y = (int)(r % d.Hd);
        const int b = (int)(r / d.Hd);
        const int qy = q / d.ps, qx = q - qy * d.ps;
        const int Y = y * d.ps + qy - d.pad, X = x * d.ps + qx - d.pad;
        const int uy = d.pad_mode ? min(max(Y, 0), Hup - 1) : reflect(Y, Hup), ux = d.pad_mode ? min(max(X, 0), Wup - 1) : reflect(X, Wup);
        const int sy = d.up == 2 ? uy >> 1 : uy, sx = d.up == 2 ? ux >> 1 : ux;
        const int c0 = cg * 8;
        float v[8];
        const size_t spix = ((size_t)b * d.Hs + sy) * d.Ws + sx;
        if (c0 < d.C) {
            const float4 r0 = ldg4(d.raw + spix * d.Cs + c0), r1 = ldg4(d.raw + spix * d.Cs + c0 + 4);
            const float4 a0 = ldg4(d.scale + b * d.C + c0), a1 = ldg4(d.scale + b * d.C + c0 + 4);
            const float4 h0 = ldg4(d.shift + b * d.C + c0), h1 = ldg4(d.shift + b * d.C + c0 + 4);
            v[0] = fmaf(r0.x, a0.x, h0.x); v[1] = fmaf(r0.y, a0.y, h0.y); v[2] = fmaf(r0.z, a0.z, h0.z); v[3] = fmaf(r0.w, a0.w, h0.w);
            v[4] = fmaf(r1.x, a1.x, h1.x); v[5] = fmaf(r1.y, a1.y, h1.y); v[6] = fmaf(r1.z, a1.z, h1.z); v[7] = fmaf(r1.w, a1.w, h1.w);
            if (d.act == B200SP_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
            }
            if (d.res_in) {
                const float4 s0 = ldg4(d.res_in + spix * d.C + c0), s1 = ldg4(d.res_in + spix * d.C + c0 + 4);
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
            }
            if (d.res_out && Y >= 0 && Y < d.Hs && X >= 0 && X < d.Ws) {      // interior pixel: the canonical copy
                float* ro = d.res_out + spix * d.C + c0;
                *reinterpret_cast<float4*>(ro) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(ro + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
        }
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        bf16* dst = reinterpret_cast<bf16*>(d.planes[q]) + ((((size_t)b * d.Hd + y) * d.Wd + x) * d.Cd + c0);
        *reinterpret_cast<uint4*>(dst) = o;
    }
}

__global__ void __launch_bounds__(SA_NT) in_apply_final_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, float* __restrict__ out, int B, int H,
                                                               int W, int Cs, int C) {
    const long long n = (long long)B * C * H * W;
    for (long long i = (long long)blockIdx.x * SA_NT + threadIdx.x; i < n; i += (long long)gridDim.x * SA_NT) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        const float z = fmaf(__ldg(raw + (((size_t)b * H + y) * W + x) * Cs + c), scale[b * C + c], shift[b * C + c]);
        out[i] = 1.f / (1.f + __expf(-z));
    }
}

// Row-decomposed 9x9 (k x k) convolution with very few output channels (ghiasi.py:121, 32 -> 3): the tensor-core
// pass computes, for every plane-grid pixel m', T[m'][kw*Co + co] = sum_{kh,c} in[m' + kh*Wq][c] w[co][c][kh][kw]
// (k taps instead of k*k: the operand traffic that bounds this layer drops k-fold, and N = k*Co fills the MMA
// instead of 3 of 16 columns); this kernel finishes  out[b,h,w,co] = sum_kw T[(b,h,w+kw)][kw*Co + co]  and
// accumulates the InstanceNorm sums.  One thread per output pixel; T rows are re-read from L1/L2.
__global__ void __launch_bounds__(SA_NT) kwsum_kernel(const float* __restrict__ T, float* __restrict__ out, float* __restrict__ stats,
                                                     int B, int Ho, int Wo, int Wq, int Nt, int k, int Co, int No, int Np) {
    __shared__ float s_sum[8], s_sq[8];
    const int b = blockIdx.y;
    if (threadIdx.x < 8) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    __syncthreads();
    float ls[4] = {0.f, 0.f, 0.f, 0.f}, lq[4] = {0.f, 0.f, 0.f, 0.f};
    const int npix = Ho * Wo;
    for (int i = blockIdx.x * SA_NT + threadIdx.x; i < npix; i += gridDim.x * SA_NT) {
        const int h = i / Wo, w = i - h * Wo;
        const float* t = T + ((size_t)(b * Ho + h) * Wq + w) * Nt;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kw = 0; kw < k; ++kw)
            for (int co = 0; co < Co; ++co) acc[co] += __ldg(t + (size_t)kw * Nt + kw * Co + co);
        float* o = out + ((size_t)b * npix + i) * No;
        for (int co = 0; co < No; ++co) o[co] = co < Co ? acc[co] : 0.f;
        for (int co = 0; co < Co; ++co) { ls[co] += acc[co]; lq[co] = fmaf(acc[co], acc[co], lq[co]); }
    }
    for (int co = 0; co < Co; ++co) {
        float a = ls[co], q = lq[co];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_sum[co], a); atomicAdd(&s_sq[co], q); }
    }
    __syncthreads();
    if (threadIdx.x < Co) {
        atomicAdd(stats + (size_t)b * 2 * Np + threadIdx.x, s_sum[threadIdx.x]);
        atomicAdd(stats + (size_t)b * 2 * Np + Np + threadIdx.x, s_sq[threadIdx.x]);
    }
}

__global__ void __launch_bounds__(SA_NT) style_embed_kernel(const float* __restrict__ noise, const float* __restrict__ A,
                                                            const float* __restrict__ mean, const float* __restrict__ base, float alpha,
                                                            float* __restrict__ out, int B, int D) {
    const int i = blockIdx.x * SA_NT + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, j = i - b * D;
    float a = 0.f;
    for (int k = 0; k < D; ++k) a = fmaf(noise[b * D + k], A[j * D + k], a);     // (noise A^T)[b][j]
    out[i] = alpha * (a + mean[j]) + (1.f - alpha) * base[j];
}

__global__ void __launch_bounds__(SA_NT) style_linear_kernel(const float* __restrict__ emb, const float* __restrict__ Wt,
                                                             const float* __restrict__ bias, float* __restrict__ out, int B, int D, int T) {
    const int i = blockIdx.x * SA_NT + threadIdx.x;
    if (i >= B * T) return;
    const int b = i / T, t = i - b * T;
    float a = bias[t];
    for (int k = 0; k < D; ++k) a = fmaf(emb[b * D + k], __ldg(Wt + (size_t)t * D + k), a);
    out[i] = a;
}

inline int sa_grid(long long n) {
    long long g = (n + SA_NT - 1) / SA_NT;
    const long long cap = (long long)NUM_SMS * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

extern "C" int b200sp_sa_prep(const float* x_nchw, void* plane, int B, int H, int W, int pad, int Cd, void* stream) {
    if (Cd < 3 || pad >= H || pad >= W) return B200SP_EINVAL;
    const long long n = (long long)B * (H + 2 * pad) * (W + 2 * pad);
    sa_prep_kernel<<<sa_grid(n), SA_NT, 0, (cudaStream_t)stream>>>(x_nchw, (bf16*)plane, B, H, W, pad, Cd);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_in_finalize(float* stats, const float* gamma, const float* beta, int gb_stride, float* scale, float* shift,
                                  int B, int C, int N_pad, int HW, float eps, void* stream) {
    in_finalize_kernel<<<ceil_div((long long)B * N_pad, SA_NT), SA_NT, 0, (cudaStream_t)stream>>>(stats, gamma, beta, gb_stride, scale, shift,
                                                                                              B, C, N_pad, 1.f / (float)HW, eps);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_in_apply(const b200sp_in_apply_desc* d, void* stream) {
    if (!d || d->Cd % 8 || d->C % 8 || d->Cs % 4 || (d->ps != 1 && d->ps != 2) || (d->up != 1 && d->up != 2)) return B200SP_EINVAL;
    if (d->res_out && (d->ps != 1 || d->up != 1)) return B200SP_EINVAL;
    if (d->pad >= d->Hs * d->up || d->pad >= d->Ws * d->up) return B200SP_EINVAL;
    const long long n = (long long)d->B * d->Hd * d->Wd * (d->Cd / 8) * d->ps * d->ps;
    in_apply_kernel<<<sa_grid(n), SA_NT, 0, (cudaStream_t)stream>>>(*d);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_in_apply_final(const float* raw, const float* scale, const float* shift, float* out_nchw,
                                     int B, int H, int W, int Cs, int C, void* stream) {
    in_apply_final_kernel<<<sa_grid((long long)B * C * H * W), SA_NT, 0, (cudaStream_t)stream>>>(raw, scale, shift, out_nchw, B, H, W, Cs, C);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_style_embed(const float* noise, const float* A, const float* mean, const float* base, float alpha,
                                  float* out, int B, int D, void* stream) {
    style_embed_kernel<<<ceil_div((long long)B * D, SA_NT), SA_NT, 0, (cudaStream_t)stream>>>(noise, A, mean, base, alpha, out, B, D);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_style_linear(const float* emb, const float* W, const float* bias, float* out, int B, int D, int T, void* stream) {
    style_linear_kernel<<<ceil_div((long long)B * T, SA_NT), SA_NT, 0, (cudaStream_t)stream>>>(emb, W, bias, out, B, D, T);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_conv_kwsum(const float* T, float* out, float* stats, int B, int Ho, int Wo, int Wq, int Nt, int k, int Co, int N_out,
                                 int N_pad, void* stream) {
    if (Co > 4 || N_out > 4 || k * Co > Nt) return B200SP_EINVAL;
    int gx = ceil_div((long long)Ho * Wo, SA_NT);
    if (gx > 64) gx = 64;
    kwsum_kernel<<<dim3(gx, B), SA_NT, 0, (cudaStream_t)stream>>>(T, out, stats, B, Ho, Wo, Wq, Nt, k, Co, N_out, N_pad);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
