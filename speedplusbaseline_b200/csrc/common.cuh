// common.cuh -- shared device helpers for libb200sp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <utility>
#include "b200sp.h"

extern long long g_b200sp_launches;      // defined in misc.cu
#define B200SP_COUNT_LAUNCH() (++g_b200sp_launches)
#define B200SP_RETURN_LAST()  do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

#define NUM_SMS 148

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// A kernel launched with b200sp_launch_pdl() may start (block scheduling, barrier / TMEM / tensor-map set-up) while its
// predecessor in the stream is still draining; it must execute pdl_wait() before it touches any global memory, and it calls
// pdl_trigger() early so that ITS successor can do the same.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool b200sp_pdl_enabled();               // misc.cu: env B200SP_PDL (default on)
template <typename... KArgs, typename... Args>
static inline cudaError_t b200sp_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = b200sp_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

typedef __nv_bfloat16 bf16;

// Branch-free activations.  Every supported activation except sigmoid is
//   act(z) = min(max(z,0) + slope*min(z,0), hi)      relu: (0, inf)  relu6: (0, 6)  leaky: (0.2, inf)  none: (1, inf)
// so the per-element cost is 4 instructions with two kernel-uniform constants instead of a switch.
struct ActP { float slope, hi; };
__device__ __forceinline__ ActP act_params(int act) {
    ActP a;
    a.slope = act == B200SP_ACT_NONE ? 1.f : (act == B200SP_ACT_LEAKY02 ? 0.2f : 0.f);
    a.hi = act == B200SP_ACT_RELU6 ? 6.f : __int_as_float(0x7f800000);
    return a;
}
__device__ __forceinline__ float act_fwd(float z, ActP a) {
    return fminf(fmaf(a.slope, fminf(z, 0.f), fmaxf(z, 0.f)), a.hi);
}
// derivative w.r.t. the pre-activation z (hardtanh backward: open interval)
__device__ __forceinline__ float act_bwd(float z, ActP a) {
    const float d = z > 0.f ? 1.f : a.slope;
    return z < a.hi ? d : 0.f;
}
__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == B200SP_ACT_SIGMOID) return 1.f / (1.f + __expf(-z));
    return act_fwd(z, act_params(act));
}
__device__ __forceinline__ float act_bwd(float z, int act) { return act_bwd(z, act_params(act)); }

// ---- 4-element vector access for float / bf16 storage -------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ float4 ld_plain(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
    static __device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
    static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ float4 ld(const bf16* p) {
        uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ float4 ld_plain(const bf16* p) { return ld(p); }
    static __device__ __forceinline__ void st(bf16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
    static __device__ __forceinline__ void st2(bf16* p, float a, float b) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
    }
    static __device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// apply the virtual-tensor transform to 4 consecutive channels starting at c
__device__ __forceinline__ float4 vt_apply4(const b200sp_vtensor& t, float4 x, float4 x2, int c) {
    if (t.mode == B200SP_VT_PLAIN) return x;
    float4 a = ldg4(t.p0 + c), b = ldg4(t.p1 + c);
    if (t.mode == B200SP_VT_BNACT) {
        const ActP ap = act_params(t.act);
        return make_float4(act_fwd(fmaf(x.x, a.x, b.x), ap), act_fwd(fmaf(x.y, a.y, b.y), ap),
                           act_fwd(fmaf(x.z, a.z, b.z), ap), act_fwd(fmaf(x.w, a.w, b.w), ap));
    }
    float4 d = ldg4(t.p2 + c);
    return make_float4(fmaf(a.x, x.x, fmaf(b.x, x2.x, d.x)), fmaf(a.y, x.y, fmaf(b.y, x2.y, d.y)),
                       fmaf(a.z, x.z, fmaf(b.z, x2.z, d.z)), fmaf(a.w, x.w, fmaf(b.w, x2.w, d.w)));
}

// load 4 consecutive channels of a virtual tensor at flat element offset `off` (channel index c)
template <typename T>
__device__ __forceinline__ float4 vt_load4(const b200sp_vtensor& t, size_t off, int c) {
    float4 x = Vec4<T>::ld(reinterpret_cast<const T*>(t.x) + off);
    float4 x2 = f4zero();
    if (t.mode == B200SP_VT_DY) x2 = Vec4<T>::ld(reinterpret_cast<const T*>(t.x2) + off);
    return vt_apply4(t, x, x2, c);
}

__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---- BatchNorm "last CTA" finalisers --------------------------------------------------------
// forward: sums -> scale/shift/mean/rstd, running stats; zeroes the accumulators.
__device__ __forceinline__ void bn_fwd_finalize_channel(const b200sp_bnfwd& bn, int c, double count) {
    double s = ld_cg_f64(bn.sum + c), q = ld_cg_f64(bn.sumsq + c);
    double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    double rstd = 1.0 / sqrt(var + (double)bn.eps);
    float sc = (float)((double)bn.gamma[c] * rstd);
    bn.scale[c] = sc;
    bn.shift[c] = (float)((double)bn.beta[c] - mean * (double)bn.gamma[c] * rstd);
    bn.mean[c] = (float)mean;
    bn.rstd[c] = (float)rstd;
    if (bn.running_mean) {
        double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        bn.running_mean[c] = (float)((1.0 - bn.momentum) * (double)bn.running_mean[c] + bn.momentum * mean);
        bn.running_var[c] = (float)((1.0 - bn.momentum) * (double)bn.running_var[c] + bn.momentum * unb);
    }
    bn.sum[c] = 0.0;
    bn.sumsq[c] = 0.0;
}
// backward: s1 = sum g, s2 = sum g*xhat  ->  dgamma, dbeta, dy = cA*g + cB*y + cC
__device__ __forceinline__ void bn_bwd_finalize_channel(const b200sp_bnbwd& bn, int c, double count) {
    double s1 = ld_cg_f64(bn.s1 + c), s2 = ld_cg_f64(bn.s2 + c);
    double sc = bn.scale[c], rstd = bn.rstd[c], mean = bn.mean[c];
    double cB = -sc * rstd * s2 / count;
    bn.cA[c] = (float)sc;
    bn.cB[c] = (float)cB;
    bn.cC[c] = (float)(-sc * s1 / count - cB * mean);
    bn.dgamma[c] += (float)s2;
    bn.dbeta[c] += (float)s1;
    bn.s1[c] = 0.0;
    bn.s2[c] = 0.0;
}

// All channels by the threads of one CTA.  Four channels per thread at a time with EVERY load of the batch issued before the first
// store: the per-channel form above costs ~4 dependent L2 round trips per channel (the stores may alias the later loads, so the
// compiler cannot hoist them), which made the last CTA's 10 iterations over 1280 channels a 20-25 us serial tail of every
// small-map depthwise launch (19 ns per channel in the r3t per-launch profile).
#ifndef BN_FIN_NB
#define BN_FIN_NB 4          // measured on the whole step: 2 -> see profiles/r2_step_switches.txt, 4 -> 5.11 ms, 8 -> 5.25 ms (register pressure where it is inlined)
#endif
__device__ __forceinline__ void bn_fwd_finalize_all(const b200sp_bnfwd& bn, int C, double count, int tid, int nthreads) {
    constexpr int NB = BN_FIN_NB;
    for (int c0 = tid; c0 < C; c0 += nthreads * NB) {
        double s[NB], q[NB];
        float ga[NB], be[NB], rm[NB], rv[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = c0 + j * nthreads;
            const bool ok = c < C;
            s[j] = ok ? ld_cg_f64(bn.sum + c) : 0.0;
            q[j] = ok ? ld_cg_f64(bn.sumsq + c) : 0.0;
            ga[j] = ok ? bn.gamma[c] : 0.f;
            be[j] = ok ? bn.beta[c] : 0.f;
            rm[j] = (ok && bn.running_mean) ? bn.running_mean[c] : 0.f;
            rv[j] = (ok && bn.running_mean) ? bn.running_var[c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = c0 + j * nthreads;
            if (c >= C) continue;
            const double mean = s[j] / count;
            double var = q[j] / count - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + (double)bn.eps);
            bn.scale[c] = (float)((double)ga[j] * rstd);
            bn.shift[c] = (float)((double)be[j] - mean * (double)ga[j] * rstd);
            bn.mean[c] = (float)mean;
            bn.rstd[c] = (float)rstd;
            if (bn.running_mean) {
                const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
                bn.running_mean[c] = (float)((1.0 - bn.momentum) * (double)rm[j] + bn.momentum * mean);
                bn.running_var[c] = (float)((1.0 - bn.momentum) * (double)rv[j] + bn.momentum * unb);
            }
            bn.sum[c] = 0.0;
            bn.sumsq[c] = 0.0;
        }
    }
}
__device__ __forceinline__ void bn_bwd_finalize_all(const b200sp_bnbwd& bn, int C, double count, int tid, int nthreads) {
    constexpr int NB = BN_FIN_NB;
    for (int c0 = tid; c0 < C; c0 += nthreads * NB) {
        double s1[NB], s2[NB];
        float sc[NB], rs[NB], mu[NB], dg[NB], db[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = c0 + j * nthreads;
            const bool ok = c < C;
            s1[j] = ok ? ld_cg_f64(bn.s1 + c) : 0.0;
            s2[j] = ok ? ld_cg_f64(bn.s2 + c) : 0.0;
            sc[j] = ok ? bn.scale[c] : 0.f;
            rs[j] = ok ? bn.rstd[c] : 0.f;
            mu[j] = ok ? bn.mean[c] : 0.f;
            dg[j] = ok ? bn.dgamma[c] : 0.f;
            db[j] = ok ? bn.dbeta[c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int c = c0 + j * nthreads;
            if (c >= C) continue;
            const double scd = sc[j], rstd = rs[j], mean = mu[j];
            const double cB = -scd * rstd * s2[j] / count;
            bn.cA[c] = (float)scd;
            bn.cB[c] = (float)cB;
            bn.cC[c] = (float)(-scd * s1[j] / count - cB * mean);
            bn.dgamma[c] = dg[j] + (float)s2[j];
            bn.dbeta[c] = db[j] + (float)s1[j];
            bn.s1[c] = 0.0;
            bn.s2[c] = 0.0;
        }
    }
}

// Elect the last CTA of the grid.  Call after this CTA's global atomics.  Returns true in every
// thread of the last CTA (after which it may read the accumulators with ld.cg).
__device__ __forceinline__ bool grid_last_cta(uint32_t* ticket, uint32_t total) {
    __shared__ int s_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = atomicAdd(ticket, 1u);
        s_is_last = (t == total - 1u);
        if (s_is_last) *ticket = 0u;
    }
    __syncthreads();
    bool last = s_is_last != 0;
    if (last) __threadfence();
    return last;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
