// optim.cu -- global-norm clip + AdamW over ONE flat fp32 parameter buffer.
// Replaces torch.optim.AdamW (src/nets/build.py:72-74) and clip_grad_norm_/clip_grad_value_
// (src/core/trainer.py:90,97,177,184; src/core/dann.py:99): ~10 tiny kernels x 176 tensors in the
// eager reference become two HBM-bound launches (28 B/param + 4 B/param for the norm pass).
// Hyper-parameters and the step counter live in device memory so the step replays inside a CUDA graph.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n4, long long n, b200sp_adamw_hp* hp) {
    __shared__ float s[8];
    float a = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = ldg4(g + i * 4);
        a = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, a))));
    }
    if (blockIdx.x == 0) for (long long i = n4 * 4 + threadIdx.x; i < n; i += 256) a = fmaf(g[i], g[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += (double)s[i];
        atomicAdd(&hp->sqnorm, t);
    }
}

// torch/optim/adam.py _single_tensor_adam (decoupled decay); torch/nn/utils/clip_grad.py
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ plow, long long n4, long long n,
                                                    const b200sp_adamw_hp* __restrict__ hp) {
    const float lr = hp->lr, b1 = hp->beta1, b2 = hp->beta2, eps = hp->eps, wd = hp->weight_decay;
    const int step = hp->step + 1;
    const float gs = hp->grad_scale;
    float coef = gs;
    if (hp->clip_mode == 1) {
        const float total = sqrtf((float)hp->sqnorm) * gs;
        coef = gs * fminf(hp->max_norm / (total + 1e-6f), 1.0f);
    }
    const float cv = hp->clip_mode == 2 ? hp->clip_value : 3.0e38f;
    const float bc1 = 1.f - powf(b1, (float)step);
    const float bc2 = 1.f - powf(b2, (float)step);
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
    const float decay = 1.f - lr * wd;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg = fminf(fmaxf(gg * coef, -cv), cv);
        pp *= decay;
        mm = mm + (gg - mm) * (1.f - b1);                 // lerp_
        vv = fmaf(vv, b2, (1.f - b2) * gg * gg);
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp = pp - step_size * (mm / denom);
    };
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 pp = *reinterpret_cast<float4*>(p + i * 4), mm = *reinterpret_cast<float4*>(m + i * 4),
               vv = *reinterpret_cast<float4*>(v + i * 4);
        const float4 gg = ldg4(g + i * 4);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(p + i * 4) = pp;
        *reinterpret_cast<float4*>(m + i * 4) = mm;
        *reinterpret_cast<float4*>(v + i * 4) = vv;
        if (plow) Vec4<bf16>::st(plow + i * 4, pp);
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += 256) {
            upd(p[i], g[i], m[i], v[i]);
            if (plow) plow[i] = __float2bfloat16_rn(p[i]);
        }
}

// runs after adamw_kernel in stream order: publish the norm, bump the step, clear the scratch
__global__ void adamw_post_kernel(b200sp_adamw_hp* hp) {
    hp->last_norm = sqrtf((float)hp->sqnorm) * hp->grad_scale;
    hp->sqnorm = 0.0;
    hp->step += 1;
}

}  // namespace

extern "C" int b200sp_grad_sqnorm(const float* g, int64_t n, b200sp_adamw_hp* hp, void* stream) {
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    if (grid > NUM_SMS * 4) grid = NUM_SMS * 4;
    if (grid < 1) grid = 1;
    sqnorm_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(g, n4, n, hp);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_adamw_step(float* p, const float* g, float* m, float* v, void* p_lowp,
                                 int64_t n, b200sp_adamw_hp* hp, void* stream) {
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    if (grid > NUM_SMS * 8) grid = NUM_SMS * 8;
    if (grid < 1) grid = 1;
    adamw_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_lowp, n4, n, hp);
    B200SP_COUNT_LAUNCH();
    adamw_post_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hp);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
