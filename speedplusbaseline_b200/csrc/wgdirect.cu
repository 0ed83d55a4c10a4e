// wgdirect.cu -- weight gradient of the long-M / tiny-N*K 1x1 convolutions as a streaming fp32 reduction on the CUDA cores.
//   dW[n][k] += sum_m dy(m, n) * xh(m, k)      dy = cA[n]*g + cB[n]*y + cC[n] (folded BatchNorm backward) | g
//                                                xh = act(sc[k]*x + sh[k])       (BatchNorm + activation on load) | x
// For the first MobileNetV2 stages (park2019.py:100-108 / torchvision mobilenetv2.py:42-62; M = 48*112*112 .. 48*28*28, N*K <= 6144)
// the tensor-core kernel is bound by its per-k-block operand hand-off: 32 rows of M per ~1800 cycles (r3i timeline of
// [602112,16,32]: 103 us for 154 MB = 1.5 TB/s), while the arithmetic is 0.2-0.9 GFMA -- a few microseconds of FFMA2.  Here a CTA
// streams R-row slabs (contiguous in memory: ~24 KB) of the three raw tensors into shared memory with 1-d bulk copies
// (cp.async.bulk + mbarrier, 3 stages: three instructions per slab instead of ~1500 per-thread cp.async), applies the per-channel transforms in
// place, and every thread accumulates an 8 x 8 register tile of dW (packed fp32x2 FMAs: 32 per staged row for 4 LDS.128), row
// groups working on different rows of the slab; partial tiles meet in shared memory and leave as one vector red.add per 4 weights.
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WG_NT = 256, WG_ST = 3;
enum { WG_PLAIN = 0, WG_XF = 1 };

struct WgArgs {
    const float *g, *y, *x;                 // [M][N], [M][N] (DY only), [M][K]
    const float *cA, *cB, *cC, *sc, *sh;    // per-channel parameters
    float* dw;                              // [N][K], accumulated
    long long MN, MK;                       // element counts (row validity of a contiguous slab = element index below these)
    int N, K, act;
    int G, rpg, R;                          // row groups, rows per group and slab, rows per slab = G * rpg
    int nslab, slab_per_cta;
};

// contiguous global -> shared bulk copy (bytes: multiple of 16), completion counted on an mbarrier
__device__ __forceinline__ void wg_bulk(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(WG_NT, 2) wgdirect_kernel(const WgArgs a) {
    extern __shared__ __align__(16) float ws[];
    const int N = a.N, K = a.K, R = a.R;
    const int gN = R * N, gK = R * K;                             // floats per slab of g / y and of x
    const int stage_f = (AMODE == WG_XF ? 2 * gN : gN) + gK;      // one stage: g | [y] | x
    float* s_par = ws + WG_ST * stage_f;                          // cA cB cC [N] | sc sh [K]
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (AMODE == WG_XF) for (int i = tid; i < N; i += WG_NT) { s_par[i] = __ldg(a.cA + i); s_par[N + i] = __ldg(a.cB + i); s_par[2 * N + i] = __ldg(a.cC + i); }
    if (BMODE == WG_XF) for (int i = tid; i < K; i += WG_NT) { s_par[3 * N + i] = __ldg(a.sc + i); s_par[3 * N + K + i] = __ldg(a.sh + i); }
    const ActP act = act_params(a.act);
    const int s0 = blockIdx.x * a.slab_per_cta, s1 = min(a.nslab, s0 + a.slab_per_cta);

    uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + 3 * N + 2 * K);      // [WG_ST] slab landed
    if (tid == 0) {
        for (int i = 0; i < WG_ST; ++i) tc::mbar_init(&bars[i], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    // one thread issues the slab: the valid part of each block is one contiguous range (a slab starts on a row boundary)
    auto issue = [&](int slab, int buf) {
        float* sg = ws + buf * stage_f;
        float* sx = sg + (AMODE == WG_XF ? 2 * gN : gN);
        const long long e0 = (long long)slab * gN, f0 = (long long)slab * gK;
        const uint32_t nb = (uint32_t)(min((long long)gN, a.MN - e0) * 4), kb = (uint32_t)(min((long long)gK, a.MK - f0) * 4);
        tc::fence_proxy_async_smem();                     // the buffer was last touched through the generic proxy (transform / reads)
        tc::mbar_arrive_expect_tx(&bars[buf], nb * (AMODE == WG_XF ? 2u : 1u) + kb);
        wg_bulk(sg, a.g + e0, nb, &bars[buf]);
        if (AMODE == WG_XF) wg_bulk(sg + gN, a.y + e0, nb, &bars[buf]);
        wg_bulk(sx, a.x + f0, kb, &bars[buf]);
    };
    if (tid == 0)
        for (int s = 0; s < WG_ST - 1; ++s)
            if (s0 + s < s1) issue(s0 + s, s);
    // thread -> (row group, 8 x 8 tile): n in {4 tn .. +3} u {N/2 + 4 tn .. +3}, k likewise (two 16-byte reads each, consecutive
    // lanes on consecutive 16-byte pieces)
    const int tkn = K >> 3, tpg = (N >> 3) * tkn;
    const int rg = tid / tpg, t = tid - rg * tpg;
    const int tn = t / tkn, tk = t - tn * tkn;
    const bool worker = rg < a.G;
    const int nA = 4 * tn, nB = (N >> 1) + 4 * tn, kA = 4 * tk, kB = (K >> 1) + 4 * tk;
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    // channel of this thread's first piece in the transform passes; a pass advances by 1024 floats
    const int stepN = (WG_NT * 4) % N, stepK = (WG_NT * 4) % K;
    const int n_first = (tid * 4) % N, k_first = (tid * 4) % K;

    for (int slab = s0; slab < s1; ++slab) {
        const int it = slab - s0, buf = it % WG_ST;
        const uint32_t par = (uint32_t)(it / WG_ST) & 1u;
        if (tid < 32) { while (!tc::mbar_try_wait_hint(&bars[buf], par, 100000u)) {} }
        __syncthreads();                                   // slab landed; the buffer refilled below was read two iterations ago by everyone
        tc::mbar_try_wait(&bars[buf], par);                // completes at once: every thread observes the bulk-copy phase itself
        if (tid == 0 && slab + WG_ST - 1 < s1) issue(slab + WG_ST - 1, (it + WG_ST - 1) % WG_ST);
        float* sg = ws + buf * stage_f;
        float* sx = sg + (AMODE == WG_XF ? 2 * gN : gN);
        // ---- transforms in place (rows past M were zero-filled and must stay zero) ----
        if (AMODE == WG_XF) {
            const long long e0 = (long long)slab * gN;
            int n = n_first;
            for (int i = tid * 4; i < gN; i += WG_NT * 4) {
                float4 v = f4zero();
                if (e0 + i < a.MN) {
                    const float4 gq = *reinterpret_cast<const float4*>(sg + i), yq = *reinterpret_cast<const float4*>(sg + gN + i);
                    const float4 pa = *reinterpret_cast<const float4*>(s_par + n), pb = *reinterpret_cast<const float4*>(s_par + N + n),
                                 pc = *reinterpret_cast<const float4*>(s_par + 2 * N + n);
                    v = make_float4(fmaf(pa.x, gq.x, fmaf(pb.x, yq.x, pc.x)), fmaf(pa.y, gq.y, fmaf(pb.y, yq.y, pc.y)),
                                    fmaf(pa.z, gq.z, fmaf(pb.z, yq.z, pc.z)), fmaf(pa.w, gq.w, fmaf(pb.w, yq.w, pc.w)));
                }
                *reinterpret_cast<float4*>(sg + i) = v;
                n += stepN;
                if (n >= N) n -= N;
            }
        }
        if (BMODE == WG_XF) {
            const long long f0 = (long long)slab * gK;
            int k = k_first;
            for (int i = tid * 4; i < gK; i += WG_NT * 4) {
                float4 v = f4zero();
                if (f0 + i < a.MK) {
                    const float4 xq = *reinterpret_cast<const float4*>(sx + i);
                    const float4 pa = *reinterpret_cast<const float4*>(s_par + 3 * N + k), pb = *reinterpret_cast<const float4*>(s_par + 3 * N + K + k);
                    v = make_float4(act_fwd(fmaf(xq.x, pa.x, pb.x), act), act_fwd(fmaf(xq.y, pa.y, pb.y), act),
                                    act_fwd(fmaf(xq.z, pa.z, pb.z), act), act_fwd(fmaf(xq.w, pa.w, pb.w), act));
                }
                *reinterpret_cast<float4*>(sx + i) = v;
                k += stepK;
                if (k >= K) k -= K;
            }
        }
        // a bulk copy brings only the rows that exist: the tail of the last slab is cleared here for an operand without a transform pass
        if (AMODE != WG_XF) {
            const long long e0 = (long long)slab * gN;
            if (e0 + gN > a.MN) for (int i = (int)(a.MN - e0) + tid; i < gN; i += WG_NT) sg[i] = 0.f;
        }
        if (BMODE != WG_XF) {
            const long long f0 = (long long)slab * gK;
            if (f0 + gK > a.MK) for (int i = (int)(a.MK - f0) + tid; i < gK; i += WG_NT) sx[i] = 0.f;
        }
        if (AMODE == WG_XF || BMODE == WG_XF || slab == a.nslab - 1) {
            tc::fence_proxy_async_smem();                  // these generic-proxy writes precede the bulk copy that refills the buffer
            __syncthreads();
        }
        // ---- rank-1 updates: this row group's rows of the slab ----
        if (worker) {
            const float* dr = sg + rg * a.rpg * N;
            const float* xr = sx + rg * a.rpg * K;
#pragma unroll 2
            for (int j = 0; j < a.rpg; ++j, dr += N, xr += K) {
                const float4 d0 = *reinterpret_cast<const float4*>(dr + nA), d1 = *reinterpret_cast<const float4*>(dr + nB);
                const float4 x0 = *reinterpret_cast<const float4*>(xr + kA), x1 = *reinterpret_cast<const float4*>(xr + kB);
                const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const float2 xa = make_float2(x0.x, x0.y), xb = make_float2(x0.z, x0.w), xc = make_float2(x1.x, x1.y), xd = make_float2(x1.z, x1.w);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 d2 = make_float2(dv[i], dv[i]);
                    acc[i][0] = __ffma2_rn(d2, xa, acc[i][0]);
                    acc[i][1] = __ffma2_rn(d2, xb, acc[i][1]);
                    acc[i][2] = __ffma2_rn(d2, xc, acc[i][2]);
                    acc[i][3] = __ffma2_rn(d2, xd, acc[i][3]);
                }
            }
        }
    }
    // ---- row groups meet in shared memory, one vector reduction per 4 weights leaves the CTA ----
    __syncthreads();
    float* s_out = ws;                                     // [N][K]
    for (int i = tid; i < N * K; i += WG_NT) s_out[i] = 0.f;
    __syncthreads();
    if (worker) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = (i < 4 ? nA : nB - 4) + i;
            float* o = s_out + n * K;
            atomicAdd(o + kA, acc[i][0].x); atomicAdd(o + kA + 1, acc[i][0].y); atomicAdd(o + kA + 2, acc[i][1].x); atomicAdd(o + kA + 3, acc[i][1].y);
            atomicAdd(o + kB, acc[i][2].x); atomicAdd(o + kB + 1, acc[i][2].y); atomicAdd(o + kB + 2, acc[i][3].x); atomicAdd(o + kB + 3, acc[i][3].y);
        }
    }
    __syncthreads();
    for (int i = tid * 4; i < N * K; i += WG_NT * 4) {
        const float4 v = *reinterpret_cast<const float4*>(s_out + i);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dw + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

inline int wg_mode() {          // B200SP_WGDIRECT: 0 off | 1 on (default)
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_WGDIRECT"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
inline int wg_min_m() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_WGDIRECT_MIN_M"); v = e ? atoi(e) : 65536; }
    return v;
}

template <int AMODE, int BMODE>
int wg_launch(WgArgs& a, size_t smem, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgdirect_kernel<AMODE, BMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    { cudaError_t le = b200sp_launch_pdl(wgdirect_kernel<AMODE, BMODE>, dim3(grid), dim3(WG_NT), smem, st, a); if (le != cudaSuccess) return (int)le; }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

}  // namespace

// B200SP_ENOSYS when the shape is not one this kernel takes (the caller continues with the tensor-core path)
int wgdirect_launch(const b200sp_vtensor* dy, const b200sp_vtensor* x, float* dw, int M, int N, int K, cudaStream_t st) {
    if (!wg_mode() || M < wg_min_m()) return B200SP_ENOSYS;
    if (N % 8 || K % 8 || N * K > 6144) return B200SP_ENOSYS;
    // measured in-graph against the tensor-core kernel (r3k, us): [602112,16,32] 61 / 103, [150528,24,96] 40 / 58, [150528,24,144] 55 / 87,
    // but [602112,96,16] 147 / 140 and the 37632-row layers 25-33 / 23-24: a gradient operand much wider than the activation operand
    // makes the in-place BatchNorm-backward pass over it the shared-memory bound, and short M leaves too few slabs per CTA
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("B200SP_WGDIRECT_WIDE"); wide = (e && e[0] == '1') ? 1 : 0; }
    if (!wide && N > 4 * K) return B200SP_ENOSYS;
    const int tpg = (N / 8) * (K / 8);
    if (tpg > WG_NT) return B200SP_ENOSYS;
    if (dy->mode == B200SP_VT_BNACT || x->mode == B200SP_VT_DY) return B200SP_ENOSYS;
    if (x->mode == B200SP_VT_BNACT && x->act == B200SP_ACT_SIGMOID) return B200SP_ENOSYS;
    if (((uintptr_t)dy->x | (uintptr_t)dy->x2 | (uintptr_t)x->x | (uintptr_t)dw) & 15) return B200SP_ENOSYS;
    WgArgs a = {};
    a.g = reinterpret_cast<const float*>(dy->x); a.y = reinterpret_cast<const float*>(dy->x2); a.x = reinterpret_cast<const float*>(x->x);
    a.cA = dy->p0; a.cB = dy->p1; a.cC = dy->p2; a.sc = x->p0; a.sh = x->p1;
    a.dw = dw; a.N = N; a.K = K; a.act = x->act;
    a.MN = (long long)M * N; a.MK = (long long)M * K;
    const bool dyx = dy->mode == B200SP_VT_DY, xx = x->mode == B200SP_VT_BNACT;
    const int rowf = (dyx ? 2 : 1) * N + K;
    a.G = WG_NT / tpg;
    a.rpg = (24 * 1024) / (a.G * rowf * 4);               // ~24 KB per slab: 2 slabs x 2 CTAs in flight per SM cover the HBM latency
    if (a.rpg < 1) a.rpg = 1;
    a.R = a.G * a.rpg;
    const size_t stage_f = (size_t)a.R * rowf;
    size_t smem = sizeof(float) * (WG_ST * stage_f + 3 * N + 2 * K) + 64;
    if (smem < sizeof(float) * (size_t)N * K) smem = sizeof(float) * (size_t)N * K;
    if (smem > 110 * 1024) return B200SP_ENOSYS;
    a.nslab = ceil_div(M, a.R);
    int grid = 2 * NUM_SMS;
    if (grid > a.nslab) grid = a.nslab;
    a.slab_per_cta = ceil_div(a.nslab, grid);
    grid = ceil_div(a.nslab, a.slab_per_cta);
    if (dyx && xx) return wg_launch<WG_XF, WG_XF>(a, smem, grid, st);
    if (dyx) return wg_launch<WG_XF, WG_PLAIN>(a, smem, grid, st);
    if (xx) return wg_launch<WG_PLAIN, WG_XF>(a, smem, grid, st);
    return wg_launch<WG_PLAIN, WG_PLAIN>(a, smem, grid, st);
}
