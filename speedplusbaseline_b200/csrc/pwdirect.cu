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

// ---- data gradient: dX[M,KO] = dY[M,NI] * W[NI,KO]  (+ skip) -> g = dX * act'(z_in), BN-backward sums of the input's BN ----
// Same structure as pwf_kernel with the roles swapped: the "input row" is the (virtual) gradient row dy[p][0..NI) -- BatchNorm
// backward folded into the load, dy = cA*g + cB*y + cC -- and the weight panel W[NI][KO] is already [in][out].  The epilogue is
// the b200sp_pw_dgrad contract (tcgemm.cu TCG_EPI_DGRAD): v = acc*scale_out (+ skip); with `bn`: v *= act'(bn.y*scale+shift),
// s1 += v, s2 += v*(bn.y-mean)*rstd (two column-sum tiles per warp), last CTA -> bn_bwd_finalize_channel.
template <int NI, int KO, int NC, int DM>
__global__ void __launch_bounds__(PW_NT, NI <= 32 ? 1 : 2) pwd_kernel(const b200sp_vtensor dy, const float* __restrict__ w,
                                                                       const float* __restrict__ skip, const float scale_out,
                                                                       float* __restrict__ gout, const b200sp_bnbwd bn,
                                                                       const int has_bn, const int M, const double count) {
    static_assert(NI % 8 == 0 && KO % NC == 0 && NC % 4 == 0 && NC <= 32, "shape table");
    constexpr bool XREG = NI <= 32;
    constexpr int NCH = KO / NC;
    static_assert(XREG || NCH == 1, "streaming the reduction needs all outputs in one chunk");
    extern __shared__ __align__(16) float pw_smem[];
    float* s_w = pw_smem;                          // [NI][KO]
    float* s_cA = s_w + NI * KO;                   // [NI] x3: dy coefficients (DM == 2)
    float* s_cB = s_cA + NI;
    float* s_cC = s_cB + NI;
    float* s_sc = s_cC + NI;                       // [KO] x4: scale, shift, mean, rstd of the input's BN
    float* s_sh = s_sc + KO;
    float* s_mu = s_sh + KO;
    float* s_rs = s_mu + KO;
    float* s_t = s_rs + KO;                        // [PW_WARPS][2][32][NC + 1]
    float* s_red = s_t + PW_WARPS * 2 * 32 * (NC + 1);   // [PW_WARPS][2][KO]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_stats = has_bn && bn.s1 != nullptr;
    for (int i = tid; i < NI * KO; i += PW_NT) s_w[i] = __ldg(w + i);
    if (DM == 2) for (int i = tid; i < NI; i += PW_NT) { s_cA[i] = __ldg(dy.p0 + i); s_cB[i] = __ldg(dy.p1 + i); s_cC[i] = __ldg(dy.p2 + i); }
    for (int i = tid; i < KO; i += PW_NT) {
        s_sc[i] = (has_bn && bn.scale) ? __ldg(bn.scale + i) : 1.f;
        s_sh[i] = (has_bn && bn.scale) ? __ldg(bn.shift + i) : 0.f;
        s_mu[i] = do_stats ? __ldg(bn.mean + i) : 0.f;
        s_rs[i] = do_stats ? __ldg(bn.rstd + i) : 0.f;
    }
    __syncthreads();
    const ActP oact = act_params(bn.act);
    const float* __restrict__ gq = reinterpret_cast<const float*>(dy.x);
    const float* __restrict__ yq = reinterpret_cast<const float*>(dy.x2);
    const float* __restrict__ yb = reinterpret_cast<const float*>(bn.y);
    float* T1 = s_t + (warp * 2 + 0) * 32 * (NC + 1);
    float* T2 = s_t + (warp * 2 + 1) * 32 * (NC + 1);
    float rs1[NCH], rs2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { rs1[c] = 0.f; rs2[c] = 0.f; }

    auto dyv = [&](const float* grow, const float* yrow, int n) {      // 4 consecutive gradient channels n..n+3
        const float4 g4 = ldg4(grow + n);
        if (DM != 2) return g4;
        const float4 y4 = ldg4(yrow + n);
        const float4 a = *reinterpret_cast<const float4*>(s_cA + n), b = *reinterpret_cast<const float4*>(s_cB + n),
                     c = *reinterpret_cast<const float4*>(s_cC + n);
        return make_float4(fmaf(a.x, g4.x, fmaf(b.x, y4.x, c.x)), fmaf(a.y, g4.y, fmaf(b.y, y4.y, c.y)),
                           fmaf(a.z, g4.z, fmaf(b.z, y4.z, c.z)), fmaf(a.w, g4.w, fmaf(b.w, y4.w, c.w)));
    };
    auto finish_chunk = [&](float (&acc)[2][NC], int ch, int p0, int p1, bool v0, bool v1) {
        const float m0 = v0 ? 1.f : 0.f, m1 = v1 ? 1.f : 0.f;
        const int pc0 = min(p0, M - 1), pc1 = min(p1, M - 1);
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
            const int k = ch * NC + j;
            float4 o0 = make_float4(acc[0][j] * scale_out, acc[0][j + 1] * scale_out, acc[0][j + 2] * scale_out, acc[0][j + 3] * scale_out);
            float4 o1 = make_float4(acc[1][j] * scale_out, acc[1][j + 1] * scale_out, acc[1][j + 2] * scale_out, acc[1][j + 3] * scale_out);
            if (skip) {
                const float4 k0 = ldg4(skip + (size_t)pc0 * KO + k), k1 = ldg4(skip + (size_t)pc1 * KO + k);
                o0.x += k0.x; o0.y += k0.y; o0.z += k0.z; o0.w += k0.w;
                o1.x += k1.x; o1.y += k1.y; o1.z += k1.z; o1.w += k1.w;
            }
            float4 h0 = f4zero(), h1 = f4zero();
            if (has_bn) {
                const float4 y0 = ldg4(yb + (size_t)pc0 * KO + k), y1 = ldg4(yb + (size_t)pc1 * KO + k);
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + k), sh = *reinterpret_cast<const float4*>(s_sh + k);
                const float4 mu = *reinterpret_cast<const float4*>(s_mu + k), rs = *reinterpret_cast<const float4*>(s_rs + k);
                o0.x *= act_bwd(fmaf(y0.x, sc.x, sh.x), oact); o0.y *= act_bwd(fmaf(y0.y, sc.y, sh.y), oact);
                o0.z *= act_bwd(fmaf(y0.z, sc.z, sh.z), oact); o0.w *= act_bwd(fmaf(y0.w, sc.w, sh.w), oact);
                o1.x *= act_bwd(fmaf(y1.x, sc.x, sh.x), oact); o1.y *= act_bwd(fmaf(y1.y, sc.y, sh.y), oact);
                o1.z *= act_bwd(fmaf(y1.z, sc.z, sh.z), oact); o1.w *= act_bwd(fmaf(y1.w, sc.w, sh.w), oact);
                h0 = make_float4((y0.x - mu.x) * rs.x, (y0.y - mu.y) * rs.y, (y0.z - mu.z) * rs.z, (y0.w - mu.w) * rs.w);
                h1 = make_float4((y1.x - mu.x) * rs.x, (y1.y - mu.y) * rs.y, (y1.z - mu.z) * rs.z, (y1.w - mu.w) * rs.w);
            }
            if (v0) *reinterpret_cast<float4*>(gout + (size_t)p0 * KO + k) = o0;
            if (v1) *reinterpret_cast<float4*>(gout + (size_t)p1 * KO + k) = o1;
            if (do_stats) {
                float* t1 = T1 + lane * (NC + 1) + j;
                float* t2 = T2 + lane * (NC + 1) + j;
                t1[0] = m0 * o0.x + m1 * o1.x; t1[1] = m0 * o0.y + m1 * o1.y; t1[2] = m0 * o0.z + m1 * o1.z; t1[3] = m0 * o0.w + m1 * o1.w;
                t2[0] = m0 * o0.x * h0.x + m1 * o1.x * h1.x; t2[1] = m0 * o0.y * h0.y + m1 * o1.y * h1.y;
                t2[2] = m0 * o0.z * h0.z + m1 * o1.z * h1.z; t2[3] = m0 * o0.w * h0.w + m1 * o1.w * h1.w;
            }
        }
        if (do_stats) {
            __syncwarp();
            if (lane < NC) {
                float a = 0.f, b = 0.f;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) { a += T1[r * (NC + 1) + lane]; b += T2[r * (NC + 1) + lane]; }
                rs1[ch] += a;
                rs2[ch] += b;
            }
            __syncwarp();
        }
    };

    const int ntiles = (M + PW_TILE - 1) / PW_TILE;
    for (int t = blockIdx.x * PW_WARPS + warp; t < ntiles; t += gridDim.x * PW_WARPS) {
        const int p0 = t * PW_TILE + lane, p1 = p0 + 32;
        const bool v0 = p0 < M, v1 = p1 < M;
        const size_t r0 = (size_t)min(p0, M - 1) * NI, r1 = (size_t)min(p1, M - 1) * NI;
        if (XREG) {
            float xr[2][NI];
#pragma unroll
            for (int n = 0; n < NI; n += 4) {
                const float4 a = dyv(gq + r0, yq + r0, n), b = dyv(gq + r1, yq + r1, n);
                xr[0][n] = a.x; xr[0][n + 1] = a.y; xr[0][n + 2] = a.z; xr[0][n + 3] = a.w;
                xr[1][n] = b.x; xr[1][n + 1] = b.y; xr[1][n + 2] = b.z; xr[1][n + 3] = b.w;
            }
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                float acc[2][NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
#pragma unroll
                for (int n = 0; n < NI; ++n) {
#pragma unroll
                    for (int j = 0; j < NC; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + n * KO + ch * NC + j);
                        acc[0][j] = fmaf(xr[0][n], w4.x, acc[0][j]); acc[0][j + 1] = fmaf(xr[0][n], w4.y, acc[0][j + 1]);
                        acc[0][j + 2] = fmaf(xr[0][n], w4.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(xr[0][n], w4.w, acc[0][j + 3]);
                        acc[1][j] = fmaf(xr[1][n], w4.x, acc[1][j]); acc[1][j + 1] = fmaf(xr[1][n], w4.y, acc[1][j + 1]);
                        acc[1][j + 2] = fmaf(xr[1][n], w4.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(xr[1][n], w4.w, acc[1][j + 3]);
                    }
                }
                finish_chunk(acc, ch, p0, p1, v0, v1);
            }
        } else {
            float acc[2][NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
#pragma unroll 2
            for (int n0 = 0; n0 < NI; n0 += 8) {
                const float4 ta0 = dyv(gq + r0, yq + r0, n0), ta1 = dyv(gq + r0, yq + r0, n0 + 4);
                const float4 tb0 = dyv(gq + r1, yq + r1, n0), tb1 = dyv(gq + r1, yq + r1, n0 + 4);
                const float xa[8] = {ta0.x, ta0.y, ta0.z, ta0.w, ta1.x, ta1.y, ta1.z, ta1.w};
                const float xb[8] = {tb0.x, tb0.y, tb0.z, tb0.w, tb1.x, tb1.y, tb1.z, tb1.w};
#pragma unroll
                for (int nn = 0; nn < 8; ++nn) {
#pragma unroll
                    for (int j = 0; j < NC; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + (n0 + nn) * KO + j);
                        acc[0][j] = fmaf(xa[nn], w4.x, acc[0][j]); acc[0][j + 1] = fmaf(xa[nn], w4.y, acc[0][j + 1]);
                        acc[0][j + 2] = fmaf(xa[nn], w4.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(xa[nn], w4.w, acc[0][j + 3]);
                        acc[1][j] = fmaf(xb[nn], w4.x, acc[1][j]); acc[1][j + 1] = fmaf(xb[nn], w4.y, acc[1][j + 1]);
                        acc[1][j + 2] = fmaf(xb[nn], w4.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(xb[nn], w4.w, acc[1][j + 3]);
                    }
                }
            }
            finish_chunk(acc, 0, p0, p1, v0, v1);
        }
    }
    if (!do_stats) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        if (lane < NC) { s_red[(warp * 2 + 0) * KO + c * NC + lane] = rs1[c]; s_red[(warp * 2 + 1) * KO + c * NC + lane] = rs2[c]; }
    __syncthreads();
    for (int k = tid; k < KO; k += PW_NT) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int wv = 0; wv < PW_WARPS; ++wv) { a += s_red[(wv * 2 + 0) * KO + k]; b += s_red[(wv * 2 + 1) * KO + k]; }
        atomicAdd(bn.s1 + k, (double)a);
        atomicAdd(bn.s2 + k, (double)b);
    }
    if (grid_last_cta(bn.ticket, gridDim.x))
        for (int k = tid; k < KO; k += PW_NT) bn_bwd_finalize_channel(bn, k, count);
}

template <int NI, int KO, int NC>
int launch_dgrad_shape(const b200sp_vtensor* dy, const float* w, const float* skip, float scale_out, float* g, const b200sp_bnbwd* bn,
                       int M, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)NI * KO + 3 * NI + 4 * KO + (size_t)PW_WARPS * 2 * 32 * (NC + 1) + (size_t)PW_WARPS * 2 * KO);
    const int ntiles = (M + PW_TILE - 1) / PW_TILE;
    const int per_sm = NI <= 32 ? 1 : 2;
    int grid = (ntiles + PW_WARPS - 1) / PW_WARPS;
    if (grid > NUM_SMS * per_sm) grid = NUM_SMS * per_sm;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pwd_kernel<NI, KO, NC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pwd_kernel<NI, KO, NC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    b200sp_bnbwd b = {};
    if (bn) b = *bn;
    if (dy->mode == B200SP_VT_DY) pwd_kernel<NI, KO, NC, 2><<<grid, PW_NT, smem, st>>>(*dy, w, skip, scale_out, g, b, bn != nullptr, M, (double)M);
    else                          pwd_kernel<NI, KO, NC, 0><<<grid, PW_NT, smem, st>>>(*dy, w, skip, scale_out, g, b, bn != nullptr, M, (double)M);
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

int pwdirect_dgrad(const b200sp_vtensor* dy, const float* w, const float* skip, float scale_out, float* g, const b200sp_bnbwd* bn,
                   int M, int N, int K, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("B200SP_PWDIRECT"); on = (e && e[0] == '1') ? 1 : 0; }
    if (!on || M < 9408) return B200SP_ENOSYS;
    if (dy->mode != B200SP_VT_PLAIN && dy->mode != B200SP_VT_DY) return B200SP_ENOSYS;
    if (bn && bn->act == B200SP_ACT_SIGMOID) return B200SP_ENOSYS;
    uintptr_t al = (uintptr_t)dy->x | (uintptr_t)g | (uintptr_t)w | (uintptr_t)skip;
    if (dy->mode == B200SP_VT_DY) al |= (uintptr_t)dy->x2;
    if (bn) al |= (uintptr_t)bn->y;
    if (al & 15) return B200SP_ENOSYS;
#define PW_CASE(NI_, KO_, NC_) if (N == NI_ && K == KO_) return launch_dgrad_shape<NI_, KO_, NC_>(dy, w, skip, scale_out, g, bn, M, st);
    PW_CASE(16, 32, 32) PW_CASE(96, 16, 16) PW_CASE(24, 96, 32) PW_CASE(144, 24, 24)
    PW_CASE(24, 144, 24) PW_CASE(32, 144, 24) PW_CASE(192, 32, 32) PW_CASE(32, 192, 32)
#undef PW_CASE
    return B200SP_ENOSYS;
}
