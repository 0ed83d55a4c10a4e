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


// The other `--optimizer` choices of the reference (src/nets/build.py:63-71): torch.optim.SGD(momentum),
// RMSprop(alpha = cfg.momentum) and Adam(betas = (cfg.momentum, 0.999)), each with COUPLED (L2) weight decay
// g <- g + wd*p, after the same clip as above.  Same flat-buffer traffic pattern as adamw_kernel:
// SGD / RMSprop 20 B/param (R p,g,s1; W p,s1), Adam 28 B/param.
//   KIND 1 SGD     (torch/optim/sgd.py _single_tensor_sgd, dampening 0, no nesterov): buf = mu*buf + g; p -= lr*buf
//                  (a zero-initialised buffer reproduces torch's first-step `buf = clone(g)` exactly)
//   KIND 2 RMSprop (torch/optim/rmsprop.py _single_tensor_rmsprop, momentum 0, not centered):
//                  sq = alpha*sq + (1-alpha)*g*g; p -= lr * g / (sqrt(sq) + eps)
//   KIND 3 Adam    (torch/optim/adam.py _single_tensor_adam, coupled decay branch)
template <int KIND>
__global__ void __launch_bounds__(256) optim_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                                                    float* __restrict__ s2, bf16* __restrict__ plow, long long n4, long long n,
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
    auto upd = [&](float& pp, float gg, float& a, float& b) {
        gg = fminf(fmaxf(gg * coef, -cv), cv);
        gg = fmaf(wd, pp, gg);                                  // grad.add(param, alpha=weight_decay)
        if (KIND == 1) {
            a = fmaf(a, b1, gg);                                // buf.mul_(momentum).add_(grad)
            pp = fmaf(-lr, a, pp);                              // param.add_(buf, alpha=-lr)
        } else if (KIND == 2) {
            a = fmaf(a, b1, (1.f - b1) * gg * gg);              // square_avg.mul_(alpha).addcmul_(grad, grad, value=1-alpha)
            pp = pp - lr * (gg / (sqrtf(a) + eps));             // param.addcdiv_(grad, avg, value=-lr)
        } else {
            a = a + (gg - a) * (1.f - b1);                      // exp_avg.lerp_(grad, 1-beta1)
            b = fmaf(b, b2, (1.f - b2) * gg * gg);
            const float denom = sqrtf(b) / bc2_sqrt + eps;
            pp = pp - step_size * (a / denom);
        }
    };
    float dummy = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 pp = *reinterpret_cast<float4*>(p + i * 4), aa = *reinterpret_cast<float4*>(s1 + i * 4), bb = f4zero();
        if (KIND == 3) bb = *reinterpret_cast<float4*>(s2 + i * 4);
        const float4 gg = ldg4(g + i * 4);
        upd(pp.x, gg.x, aa.x, bb.x); upd(pp.y, gg.y, aa.y, bb.y); upd(pp.z, gg.z, aa.z, bb.z); upd(pp.w, gg.w, aa.w, bb.w);
        *reinterpret_cast<float4*>(p + i * 4) = pp;
        *reinterpret_cast<float4*>(s1 + i * 4) = aa;
        if (KIND == 3) *reinterpret_cast<float4*>(s2 + i * 4) = bb;
        if (plow) Vec4<bf16>::st(plow + i * 4, pp);
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += 256) {
            upd(p[i], g[i], s1[i], KIND == 3 ? s2[i] : dummy);
            if (plow) plow[i] = __float2bfloat16_rn(p[i]);
        }
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

extern "C" int b200sp_optim_step(int kind, float* p, const float* g, float* s1, float* s2, void* p_lowp,
                                 int64_t n, b200sp_adamw_hp* hp, void* stream) {
    if (kind == B200SP_OPT_ADAMW) return b200sp_adamw_step(p, g, s1, s2, p_lowp, n, hp, stream);
    if (!p || !g || !s1 || !hp || n < 0 || (kind == B200SP_OPT_ADAM && !s2)) return B200SP_EINVAL;
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    if (grid > NUM_SMS * 8) grid = NUM_SMS * 8;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == B200SP_OPT_SGD)          optim_kernel<1><<<(unsigned)grid, 256, 0, st>>>(p, g, s1, s2, (bf16*)p_lowp, n4, n, hp);
    else if (kind == B200SP_OPT_RMSPROP) optim_kernel<2><<<(unsigned)grid, 256, 0, st>>>(p, g, s1, s2, (bf16*)p_lowp, n4, n, hp);
    else if (kind == B200SP_OPT_ADAM)    optim_kernel<3><<<(unsigned)grid, 256, 0, st>>>(p, g, s1, s2, (bf16*)p_lowp, n4, n, hp);
    else return B200SP_EINVAL;
    B200SP_COUNT_LAUNCH();
    adamw_post_kernel<<<1, 1, 0, st>>>(hp);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
