// dann.cu -- the tail of the DANN domain classifier (src/nets/revgrad.py:75-80) and its loss
// (src/core/dann.py:85-92):  ReLU(conv1x1 320->1280 + b) is b200sp_pw_fwd with bias+ReLU in the
// epilogue; these kernels do  AvgPool2d(7) -> conv1x1 1280->1 (+b) -> binary_cross_entropy_with_logits
// (mean) forward and backward.  The gradient-reversal layer (revgrad.py:46-56) has no forward kernel
// (the clone is an alias) -- its -lambda is applied to the gradient entering the feature extractor by
// b200sp_scale_dev, from a DEVICE scalar so the step replays inside a CUDA graph while alpha changes.
#include "common.cuh"

namespace {

constexpr int DN_NT = 256;

// one CTA per image: pooled[b,c] = mean_hw h[b,hw,c];  z[b] = sum_c pooled[b,c] w3[c] + b3
template <typename T>
__global__ void __launch_bounds__(DN_NT) dann_head_fwd_kernel(const T* __restrict__ h, const float* __restrict__ w3,
                                                              const float* __restrict__ b3, float* __restrict__ pooled,
                                                              float* __restrict__ z, int HW, int C) {
    __shared__ float s_red[DN_NT / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const T* hb = h + (size_t)b * HW * C;
    const float inv = 1.f / (float)HW;
    float dot = 0.f;
    for (int c = tid * 4; c < C; c += DN_NT * 4) {
        float4 a = f4zero();
        for (int p = 0; p < HW; ++p) {
            const float4 v = Vec4<T>::ld(hb + (size_t)p * C + c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
        *reinterpret_cast<float4*>(pooled + (size_t)b * C + c) = a;
        const float4 w = ldg4(w3 + c);
        dot = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, dot))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = dot;
    __syncthreads();
    if (tid == 0) {
        float t = b3[0];
        for (int i = 0; i < DN_NT / 32; ++i) t += s_red[i];
        z[b] = t;
    }
}

// one CTA.  loss[0] = mean_b BCEWithLogits(z_b, label);  dz[b] = (sigmoid(z_b) - label)/B * loss_scale
__global__ void __launch_bounds__(DN_NT) bce_logits_kernel(const float* __restrict__ z, float label, float* __restrict__ loss,
                                                           float* __restrict__ dz, const float* __restrict__ loss_scale, int B) {
    __shared__ float s[DN_NT];
    const int tid = threadIdx.x;
    const float ls = loss_scale ? loss_scale[0] : 1.f;
    float a = 0.f;
    for (int b = tid; b < B; b += DN_NT) {
        const float x = z[b];
        a += fmaxf(x, 0.f) - x * label + log1pf(expf(-fabsf(x)));
        const float sg = 1.f / (1.f + expf(-x));
        dz[b] = (sg - label) / (float)B * ls;
    }
    s[tid] = a;
    __syncthreads();
    for (int o = DN_NT / 2; o > 0; o >>= 1) {
        if (tid < o) s[tid] += s[tid + o];
        __syncthreads();
    }
    if (tid == 0) loss[0] = s[0] / (float)B;
}

// grid (B, HW-chunks): dH[b,hw,c] = dz[b] * w3[c] / HW * (h > 0), written IN PLACE over h.
template <typename T>
__global__ void __launch_bounds__(DN_NT) dann_head_bwd_kernel(T* __restrict__ h, const float* __restrict__ dz,
                                                              const float* __restrict__ w3, int HW, int C) {
    const int b = blockIdx.x;
    const float s = dz[b] / (float)HW;
    const int C4 = C / 4;
    T* hb = h + (size_t)b * HW * C;
    for (int i = blockIdx.y * DN_NT + threadIdx.x; i < HW * C4; i += gridDim.y * DN_NT) {
        const int c = (i % C4) * 4;
        const float4 v = Vec4<T>::ld_plain(hb + (size_t)i * 4);
        const float4 w = ldg4(w3 + c);
        Vec4<T>::st(hb + (size_t)i * 4, make_float4(v.x > 0.f ? s * w.x : 0.f, v.y > 0.f ? s * w.y : 0.f,
                                                   v.z > 0.f ? s * w.z : 0.f, v.w > 0.f ? s * w.w : 0.f));
    }
}

// dw3[c] += sum_b dz[b] pooled[b,c];  db3 += sum_b dz[b]   (fixed summation order: deterministic)
__global__ void __launch_bounds__(DN_NT) dann_head_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ pooled,
                                                                float* __restrict__ dw3, float* __restrict__ db3, int B, int C) {
    const int c = blockIdx.x * DN_NT + threadIdx.x;
    if (c < C) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a = fmaf(dz[b], pooled[(size_t)b * C + c], a);
        dw3[c] += a;
    }
    if (c == 0) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dz[b];
        db3[0] += a;
    }
}

template <typename T>
__global__ void __launch_bounds__(DN_NT) scale_dev_kernel(T* __restrict__ x, long long n4, const float* __restrict__ s, float mul) {
    const float f = s[0] * mul;
    for (long long i = (long long)blockIdx.x * DN_NT + threadIdx.x; i < n4; i += (long long)gridDim.x * DN_NT) {
        float4 v = Vec4<T>::ld_plain(x + i * 4);
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        Vec4<T>::st(x + i * 4, v);
    }
}

}  // namespace

extern "C" int b200sp_dann_head_fwd(const void* h, const float* w3, const float* b3, float* pooled, float* z,
                                    int B, int HW, int C, int dtype, void* stream) {
    if (C % 4) return B200SP_EINVAL;
    if (dtype == B200SP_F32)
        dann_head_fwd_kernel<float><<<B, DN_NT, 0, (cudaStream_t)stream>>>((const float*)h, w3, b3, pooled, z, HW, C);
    else if (dtype == B200SP_BF16)
        dann_head_fwd_kernel<bf16><<<B, DN_NT, 0, (cudaStream_t)stream>>>((const bf16*)h, w3, b3, pooled, z, HW, C);
    else return B200SP_EINVAL;
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_bce_logits(const float* z, float label, float* loss, float* dz, const float* loss_scale, int B, void* stream) {
    bce_logits_kernel<<<1, DN_NT, 0, (cudaStream_t)stream>>>(z, label, loss, dz, loss_scale, B);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_dann_head_bwd(void* h_inout, const float* dz, const float* pooled, const float* w3, float* dw3, float* db3,
                                    int B, int HW, int C, int dtype, void* stream) {
    if (C % 4) return B200SP_EINVAL;
    const int chunks = ceil_div((long long)HW * (C / 4), DN_NT * 4);
    dim3 grid(B, chunks < 1 ? 1 : chunks);
    if (dtype == B200SP_F32)
        dann_head_bwd_kernel<float><<<grid, DN_NT, 0, (cudaStream_t)stream>>>((float*)h_inout, dz, w3, HW, C);
    else if (dtype == B200SP_BF16)
        dann_head_bwd_kernel<bf16><<<grid, DN_NT, 0, (cudaStream_t)stream>>>((bf16*)h_inout, dz, w3, HW, C);
    else return B200SP_EINVAL;
    B200SP_COUNT_LAUNCH();
    dann_head_wgrad_kernel<<<ceil_div(C, DN_NT), DN_NT, 0, (cudaStream_t)stream>>>(dz, pooled, dw3, db3, B, C);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_scale_dev(void* x, int64_t n, const float* s, float mul, int dtype, void* stream) {
    if (n % 4) return B200SP_EINVAL;
    const long long n4 = n / 4;
    long long g = (n4 + DN_NT - 1) / DN_NT;
    if (g > NUM_SMS * 8) g = NUM_SMS * 8;
    if (g < 1) g = 1;
    if (dtype == B200SP_F32) scale_dev_kernel<float><<<(unsigned)g, DN_NT, 0, (cudaStream_t)stream>>>((float*)x, n4, s, mul);
    else if (dtype == B200SP_BF16) scale_dev_kernel<bf16><<<(unsigned)g, DN_NT, 0, (cudaStream_t)stream>>>((bf16*)x, n4, s, mul);
    else return B200SP_EINVAL;
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
