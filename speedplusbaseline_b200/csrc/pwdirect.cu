// pwdirect.cu -- ROUND-2 CANDIDATE (env B200SP_PWDIRECT=1; never run on a GPU yet): exact-fp32 pointwise-convolution forward for
// the long-M / tiny-N*K MobileNetV2 layers (park2019.py:51 / torchvision mobilenetv2.py:42-52 with M >= 37632).
//
// Why not a GEMM: these eight shapes are HBM-bound (10-27 FLOP/B) and cost 537 us per step on the 3xTF32 tensor-core paths, whose
// producers spend 8 SASS instructions per operand ELEMENT on the fp32 -> (hi, lo) split, against 147 us of HBM time and <= 26 us
// of plain fp32 FFMA each.  Here a lane owns two pixels, the [K][N] weight panel lives in shared memory and is
// read with warp-broadcast LDS.128 (1 LDS per 8 FFMA), the input row is either held in registers (K <= 32, loop over chunks of <= 32
// output channels) or streamed in 32-byte pieces (N <= 32), BatchNorm+activation is applied on load and the BatchNorm statistics
// of the raw output are column-summed through a per-warp shared tile -- the b200sp_pw_fwd contract (raw output + bn sums + last
// CTA finalise), no split, no operand staging, no barriers in the main loop.  Results are plain fp32 (closer to torch than 3xTF32).
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int PW_NT = 256, PW_WARPS = PW_NT / 32, PW_TILE = 64;      // 64 pixels per warp-iteration: lane -> pixels lane, lane + 32

template <int K, int N, int NC, int XM>
__global__ void __launch_bounds__(PW_NT, K <= 32 ? 1 : 2) pwf_kernel(const b200sp_vtensor x, const float* __restrict__ w,
                                                                      float* __restrict__ y, const b200sp_bnfwd bn, const int M,
                                                                      const double count) {
    static_assert(K % 8 == 0 && N % NC == 0 && NC % 4 == 0 && NC <= 32, "shape table");
    constexpr bool XREG = K <= 32;                 // the two input rows of a lane live in registers
    constexpr int NCH = N / NC;
    static_assert(XREG || NCH == 1, "streaming K needs all outputs in one chunk");
    extern __shared__ __align__(16) float pw_smem[];
    float* s_w = pw_smem;                          // [K][N]   (transposed weight panel)
    float* s_sc = s_w + K * N;                     // [K] BN scale of the input (XM == 1)
    float* s_sh = s_sc + K;                        // [K]
    float* s_t = s_sh + K;                         // [PW_WARPS][32][NC + 1] column-sum tiles
    float* s_red = s_t + PW_WARPS * 32 * (NC + 1); // [PW_WARPS][2][N]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < N * K; i += PW_NT) s_w[(i % K) * N + i / K] = __ldg(w + i);
    if (XM == 1) for (int i = tid; i < K; i += PW_NT) { s_sc[i] = __ldg(x.p0 + i); s_sh[i] = __ldg(x.p1 + i); }
    __syncthreads();
    const ActP act = act_params(x.act);
    const float* __restrict__ xq = reinterpret_cast<const float*>(x.x);
    float* T = s_t + warp * 32 * (NC + 1);
    float rs[NCH], rq[NCH];                        // running column sums owned by lane < NC: channel ch*NC + lane
#pragma unroll
    for (int c = 0; c < NCH; ++c) { rs[c] = 0.f; rq[c] = 0.f; }

    auto xform = [&](float4 v, int k) {            // BN affine + activation of 4 consecutive input channels k..k+3
        if (XM == 0) return v;
        const float4 a = *reinterpret_cast<const float4*>(s_sc + k), b = *reinterpret_cast<const float4*>(s_sh + k);
        return make_float4(act_fwd(fmaf(v.x, a.x, b.x), act), act_fwd(fmaf(v.y, a.y, b.y), act),
                           act_fwd(fmaf(v.z, a.z, b.z), act), act_fwd(fmaf(v.w, a.w, b.w), act));
    };
    auto finish_chunk = [&](const float (&acc)[2][NC], int ch, int p0, int p1, bool v0, bool v1) {
        // raw output
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
            if (v0) *reinterpret_cast<float4*>(y + (size_t)p0 * N + ch * NC + j) = make_float4(acc[0][j], acc[0][j + 1], acc[0][j + 2], acc[0][j + 3]);
            if (v1) *reinterpret_cast<float4*>(y + (size_t)p1 * N + ch * NC + j) = make_float4(acc[1][j], acc[1][j + 1], acc[1][j + 2], acc[1][j + 3]);
        }
        // column sums over the 64 pixels of the tile: lane-major store, channel-major read (conflict-free with the +1 pad)
        const float m0 = v0 ? 1.f : 0.f, m1 = v1 ? 1.f : 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) T[lane * (NC + 1) + j] = m0 * acc[0][j] + m1 * acc[1][j];
        __syncwarp();
        if (lane < NC) {
            float a = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) a += T[r * (NC + 1) + lane];
            rs[ch] += a;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NC; ++j) T[lane * (NC + 1) + j] = m0 * acc[0][j] * acc[0][j] + m1 * acc[1][j] * acc[1][j];
        __syncwarp();
        if (lane < NC) {
            float a = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) a += T[r * (NC + 1) + lane];
            rq[ch] += a;
        }
        __syncwarp();
    };

    const int ntiles = (M + PW_TILE - 1) / PW_TILE;
    for (int t = blockIdx.x * PW_WARPS + warp; t < ntiles; t += gridDim.x * PW_WARPS) {
        const int p0 = t * PW_TILE + lane, p1 = p0 + 32;
        const bool v0 = p0 < M, v1 = p1 < M;
        const float* x0 = xq + (size_t)min(p0, M - 1) * K;
        const float* x1 = xq + (size_t)min(p1, M - 1) * K;
        if (XREG) {
            float xr[2][K];
#pragma unroll
            for (int k = 0; k < K; k += 4) {
                const float4 a = xform(ldg4(x0 + k), k), b = xform(ldg4(x1 + k), k);
                xr[0][k] = a.x; xr[0][k + 1] = a.y; xr[0][k + 2] = a.z; xr[0][k + 3] = a.w;
                xr[1][k] = b.x; xr[1][k + 1] = b.y; xr[1][k + 2] = b.z; xr[1][k + 3] = b.w;
            }
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                float acc[2][NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int j = 0; j < NC; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * N + ch * NC + j);
                        acc[0][j] = fmaf(xr[0][k], w4.x, acc[0][j]); acc[0][j + 1] = fmaf(xr[0][k], w4.y, acc[0][j + 1]);
                        acc[0][j + 2] = fmaf(xr[0][k], w4.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(xr[0][k], w4.w, acc[0][j + 3]);
                        acc[1][j] = fmaf(xr[1][k], w4.x, acc[1][j]); acc[1][j + 1] = fmaf(xr[1][k], w4.y, acc[1][j + 1]);
                        acc[1][j + 2] = fmaf(xr[1][k], w4.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(xr[1][k], w4.w, acc[1][j + 3]);
                    }
                }
                finish_chunk(acc, ch, p0, p1, v0, v1);
            }
        } else {
            float acc[2][NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
#pragma unroll 2
            for (int k0 = 0; k0 < K; k0 += 8) {
                const float4 a0 = ldg4(x0 + k0), a1 = ldg4(x0 + k0 + 4), b0 = ldg4(x1 + k0), b1 = ldg4(x1 + k0 + 4);
                const float4 ta0 = xform(a0, k0), ta1 = xform(a1, k0 + 4), tb0 = xform(b0, k0), tb1 = xform(b1, k0 + 4);
                const float xa[8] = {ta0.x, ta0.y, ta0.z, ta0.w, ta1.x, ta1.y, ta1.z, ta1.w};
                const float xb[8] = {tb0.x, tb0.y, tb0.z, tb0.w, tb1.x, tb1.y, tb1.z, tb1.w};
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
                    for (int j = 0; j < NC; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + (k0 + kk) * N + j);
                        acc[0][j] = fmaf(xa[kk], w4.x, acc[0][j]); acc[0][j + 1] = fmaf(xa[kk], w4.y, acc[0][j + 1]);
                        acc[0][j + 2] = fmaf(xa[kk], w4.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(xa[kk], w4.w, acc[0][j + 3]);
                        acc[1][j] = fmaf(xb[kk], w4.x, acc[1][j]); acc[1][j + 1] = fmaf(xb[kk], w4.y, acc[1][j + 1]);
                        acc[1][j + 2] = fmaf(xb[kk], w4.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(xb[kk], w4.w, acc[1][j + 3]);
                    }
                }
            }
            finish_chunk(acc, 0, p0, p1, v0, v1);
        }
    }
    // ---- CTA-level statistics: warps -> shared, one double atomic per channel per CTA, last CTA finalises ----
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        if (lane < NC) { s_red[(warp * 2 + 0) * N + c * NC + lane] = rs[c]; s_red[(warp * 2 + 1) * N + c * NC + lane] = rq[c]; }
    __syncthreads();
    for (int n = tid; n < N; n += PW_NT) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int wv = 0; wv < PW_WARPS; ++wv) { a += s_red[(wv * 2 + 0) * N + n]; b += s_red[(wv * 2 + 1) * N + n]; }
        atomicAdd(bn.sum + n, (double)a);
        atomicAdd(bn.sumsq + n, (double)b);
    }
    if (grid_last_cta(bn.ticket, gridDim.x))
        for (int n = tid; n < N; n += PW_NT) bn_fwd_finalize_channel(bn, n, count);
}

template <int K, int N, int NC>
int launch_shape(const b200sp_vtensor* x, const float* w, float* y, const b200sp_bnfwd* bn, int M, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)K * N + 2 * K + (size_t)PW_WARPS * 32 * (NC + 1) + (size_t)PW_WARPS * 2 * N);
    const int ntiles = (M + PW_TILE - 1) / PW_TILE;
    const int per_sm = K <= 32 ? 1 : 2;
    int grid = (ntiles + PW_WARPS - 1) / PW_WARPS;
    if (grid > NUM_SMS * per_sm) grid = NUM_SMS * per_sm;
    static bool attr_set = false;              // one flag per (K, N, NC) instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pwf_kernel<K, N, NC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pwf_kernel<K, N, NC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (x->mode == B200SP_VT_PLAIN) pwf_kernel<K, N, NC, 0><<<grid, PW_NT, smem, st>>>(*x, w, y, *bn, M, (double)M);
    else                            pwf_kernel<K, N, NC, 1><<<grid, PW_NT, smem, st>>>(*x, w, y, *bn, M, (double)M);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

}  // namespace

// returns B200SP_ENOSYS when the call is not one of the direct shapes (the caller then takes the GEMM path)
int pwdirect_fwd(const b200sp_vtensor* x, const float* w, const float* bias, int out_act, float* y, const b200sp_bnfwd* bn,
                 int M, int N, int K, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("B200SP_PWDIRECT"); on = (e && e[0] == '1') ? 1 : 0; }
    if (!on || bias || out_act != B200SP_ACT_NONE || !bn || M < 9408) return B200SP_ENOSYS;
    if (x->mode != B200SP_VT_PLAIN && x->mode != B200SP_VT_BNACT) return B200SP_ENOSYS;
    if (x->mode == B200SP_VT_BNACT && x->act == B200SP_ACT_SIGMOID) return B200SP_ENOSYS;
    if ((((uintptr_t)x->x | (uintptr_t)y | (uintptr_t)w) & 15) != 0) return B200SP_ENOSYS;
#define PW_CASE(K_, N_, NC_) if (K == K_ && N == N_) return launch_shape<K_, N_, NC_>(x, w, y, bn, M, st);
    PW_CASE(32, 16, 16) PW_CASE(16, 96, 32) PW_CASE(96, 24, 24) PW_CASE(24, 144, 24)
    PW_CASE(144, 24, 24) PW_CASE(144, 32, 32) PW_CASE(32, 192, 32) PW_CASE(192, 32, 32)
#undef PW_CASE
    return B200SP_ENOSYS;
}
