// opsplit.cu -- operand pre-pass of the "presplit" tcgen05 GEMM (tcgemm2.cu, PRE mode).
// The 7x7 layers of the KRN (M = 48*7*7 = 2352 rows, K / N up to 1280: park2019.py:100-118, mobilenetv2.py features 14-17) are the
// only tensor-bound 1x1 convolutions of the step, and their operands (<= 12 MB) live in L2.  For them the per-k-block converter
// chain of the general kernel (smem -> BN/activation -> tf32 hi/lo -> smem, ~1900 cycles per 32-wide k-block against ~770 of MMA
// work) is paid once per output tile COLUMN, i.e. 8-19 times per element.  This kernel applies the virtual-tensor
// transform and the tf32 split ONCE per element and writes K-major planes [hi | lo] that the GEMM's TMA unit loads straight into the
// swizzled UMMA operand tiles; an operand whose reduction dimension is the slow one in memory (weights in dgrad, both operands in
// wgrad) is transposed on the way, so all three passes run the same K-major x K-major kernel.
#include "tcgemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int OS_NT = 256;
constexpr int TR_R = 64, TR_C = 32, TR_LD = TR_R + 4;      // transposing tile: 64 source rows x 32 source columns

__device__ __forceinline__ float4 os_xform(const OpSplitJob& j, float4 x, float4 x2, int c, ActP act) {
    if (j.t.mode == B200SP_VT_PLAIN) return x;
    const float4 a = ldg4(j.t.p0 + c), b = ldg4(j.t.p1 + c);
    if (j.t.mode == B200SP_VT_BNACT) {
        // same arithmetic as the in-kernel converters (tcgemm2.cu xf_apply): one FMA, then the activation
        return make_float4(act_fwd(fmaf(x.x, a.x, b.x), act), act_fwd(fmaf(x.y, a.y, b.y), act), act_fwd(fmaf(x.z, a.z, b.z), act),
                           act_fwd(fmaf(x.w, a.w, b.w), act));
    }
    const float4 d = ldg4(j.t.p2 + c);
    return make_float4(fmaf(a.x, x.x, fmaf(b.x, x2.x, d.x)), fmaf(a.y, x.y, fmaf(b.y, x2.y, d.y)), fmaf(a.z, x.z, fmaf(b.z, x2.z, d.z)),
                       fmaf(a.w, x.w, fmaf(b.w, x2.w, d.w)));
}
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void job_plain(const OpSplitJob& j, int blk) {
    const ActP act = act_params(j.t.act);
    const int c4n = j.cols >> 2;
    const long long total = (long long)j.rows * c4n;
    const float* x = reinterpret_cast<const float*>(j.t.x);
    const float* x2 = reinterpret_cast<const float*>(j.t.x2);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long i = ((long long)blk * 4 + u) * OS_NT + threadIdx.x;
        if (i >= total) break;
        const int r = (int)(i / c4n), c = (int)(i - (long long)r * c4n) * 4;
        const size_t off = (size_t)r * j.ld + c;
        const float4 v0 = ldg4(x + off);
        const float4 v2 = j.t.mode == B200SP_VT_DY ? ldg4(x2 + off) : f4zero();
        const float4 v = os_xform(j, v0, v2, c, act);
        float* o = j.out + (size_t)r * j.out_ld + c;
        if (j.nm == 1) {
            *reinterpret_cast<float4*>(o) = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
        } else {
            float4 h, l;
            tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(o) = h;
            *reinterpret_cast<float4*>(o + j.plane) = l;
        }
    }
}

__device__ __forceinline__ void job_trans(const OpSplitJob& j, int blk, float* s) {
    const ActP act = act_params(j.t.act);
    const int tcn = (j.cols + TR_C - 1) / TR_C;
    const int tr = blk / tcn, tcx = blk - tr * tcn;
    const int r0 = tr * TR_R, c0 = tcx * TR_C;
    const float* x = reinterpret_cast<const float*>(j.t.x);
    const float* x2 = reinterpret_cast<const float*>(j.t.x2);
    // read 64 rows x 8 quads, transform, park transposed: s[c][r]
#pragma unroll
    for (int u = 0; u < TR_R * (TR_C / 4) / OS_NT; ++u) {
        const int i = u * OS_NT + threadIdx.x;
        const int r = i >> 3, c = (i & 7) * 4;
        float4 v = f4zero();
        if (r0 + r < j.rows && c0 + c < j.cols) {
            const size_t off = (size_t)(r0 + r) * j.ld + c0 + c;
            const float4 v0 = ldg4(x + off);
            const float4 v2 = j.t.mode == B200SP_VT_DY ? ldg4(x2 + off) : f4zero();
            v = os_xform(j, v0, v2, c0 + c, act);
        }
        s[(c + 0) * TR_LD + r] = v.x; s[(c + 1) * TR_LD + r] = v.y; s[(c + 2) * TR_LD + r] = v.z; s[(c + 3) * TR_LD + r] = v.w;
    }
    __syncthreads();
    // write 32 output rows (source columns) x 16 quads of source rows
#pragma unroll
    for (int u = 0; u < TR_C * (TR_R / 4) / OS_NT; ++u) {
        const int i = u * OS_NT + threadIdx.x;
        const int c = i >> 4, r = (i & 15) * 4;
        if (c0 + c >= j.cols || r0 + r >= j.rows) continue;       // rows % 4 == 0: a quad is all in or all out
        const float4 v = *reinterpret_cast<const float4*>(s + c * TR_LD + r);
        float* o = j.out + (size_t)(c0 + c) * j.out_ld + r0 + r;
        if (j.nm == 1) {
            *reinterpret_cast<float4*>(o) = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
        } else {
            float4 h, l;
            tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(o) = h;
            *reinterpret_cast<float4*>(o + j.plane) = l;
        }
    }
}

__global__ void __launch_bounds__(OS_NT) opsplit_kernel(const __grid_constant__ OpSplitJob j0, const __grid_constant__ OpSplitJob j1, int blocks0) {
    __shared__ __align__(16) float s[TR_C * TR_LD];
    // programmatic dependent launch: this grid may be scheduled while its predecessor drains, but it overwrites the workspace the
    // previous GEMM may still be reading -- nothing is touched before griddepcontrol.wait (= predecessor complete and flushed);
    // the trigger right after lets the GEMM that consumes these planes set up (barriers, TMEM) while this kernel runs
    pdl_wait();
    pdl_trigger();
    const bool first = (int)blockIdx.x < blocks0;
    const OpSplitJob& j = first ? j0 : j1;
    const int blk = first ? blockIdx.x : blockIdx.x - blocks0;
    if (j.trans) job_trans(j, blk, s);
    else job_plain(j, blk);
}

inline int job_blocks(const OpSplitJob& j) {
    if (j.trans) return ceil_div(j.rows, TR_R) * ceil_div(j.cols, TR_C);
    return ceil_div((long long)j.rows * (j.cols / 4), OS_NT * 4);
}

}  // namespace

// both operands of one GEMM in one launch
int opsplit_launch(const OpSplitJob& a, const OpSplitJob& b, cudaStream_t st) {
    const int b0 = job_blocks(a), b1 = job_blocks(b);
    { cudaError_t le = b200sp_launch_pdl(opsplit_kernel, dim3(b0 + b1), dim3(OS_NT), 0, st, a, b, b0); if (le != cudaSuccess) return (int)le; }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
