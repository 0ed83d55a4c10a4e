// fcstream.cu -- weight-streaming fully connected layers for small batches (SPN fc6..fc11, spn.py:80-99: M = batch <= 32,
// weights 16-38 M floats).  With one 128-row M tile the tensor-core GEMM is a latency chain over K (round 1: 105-125 us for the
// 151 MB fc6 weight = 1.1-1.4 TB/s); these layers are pure weight streams, so they are written as such:
//   forward  y[M][N] += x[M][K] * W[N][K]^T : CTA = 64 weight rows x a K split; W and x tiles staged with cp.async (3 stages),
//                                             thread = 4 batch rows x 4 outputs, 16-byte shared-memory reads along K;
//   dgrad    dx[M][K] += dy[M][N] * W[N][K] : CTA = 64 weight columns x an N split; dy tile transposed in shared memory so that a
//                                             thread's 4 batch values are one 16-byte read; thread = 4 batch rows x 4 columns.
// Partial sums leave through fp32 atomics (the callers pass a zeroed / to-be-accumulated output, as for the split-K GEMM).
// CUDA-core FFMA: 1.2 GFLOP per layer is ~20 us of the machine's FFMA rate, below the 23 us the 151 MB stream needs.
#include "common.cuh"

namespace {

constexpr int FC_NT = 128;                 // threads: 8 batch blocks (4 rows) x 16 output blocks (4 outputs / columns)
constexpr int FC_TN = 64;                  // weight rows (fwd) / weight columns (dgrad) per CTA
constexpr int FC_KC = 64;                  // reduction elements per stage
constexpr int FC_LD = FC_KC + 4;           // padded row pitch (floats): 16-byte reads of 8 consecutive rows cover all banks
constexpr int FC_ST = 3;                   // cp.async stages
constexpr int FC_MB = 32;                  // batch rows held (M <= 32)

__device__ __forceinline__ void cpa16(void* dst, const void* src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- forward ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FC_NT) fc_fwd_stream_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                              int M, int N, int K, int k_per_split) {
    extern __shared__ __align__(16) float fs[];
    float* s_w = fs;                                   // [FC_ST][FC_TN][FC_LD]
    float* s_x = fs + FC_ST * FC_TN * FC_LD;           // [FC_ST][FC_MB][FC_LD]
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * FC_TN;
    const int kbeg = blockIdx.y * k_per_split, kend = min(K, kbeg + k_per_split);
    const int nst = (kend - kbeg + FC_KC - 1) / FC_KC;
    auto load = [&](int st, int buf) {
        const int k0 = kbeg + st * FC_KC;
        for (int i = tid; i < FC_TN * (FC_KC / 4); i += FC_NT) {              // weights: 64 rows x 16 pieces
            const int r = i >> 4, c = (i & 15) * 4;
            const bool ok = n0 + r < N && k0 + c < kend;
            cpa16(s_w + (buf * FC_TN + r) * FC_LD + c, w + (size_t)(ok ? n0 + r : 0) * K + (ok ? k0 + c : 0), ok);
        }
        for (int i = tid; i < FC_MB * (FC_KC / 4); i += FC_NT) {              // activations: 32 rows x 16 pieces
            const int r = i >> 4, c = (i & 15) * 4;
            const bool ok = r < M && k0 + c < kend;
            cpa16(s_x + (buf * FC_MB + r) * FC_LD + c, x + (size_t)(ok ? r : 0) * K + (ok ? k0 + c : 0), ok);
        }
    };
    for (int s = 0; s < FC_ST - 1; ++s) { if (s < nst) load(s, s); cpa_commit(); }
    // lane layout: 8 batch rows x 4 outputs per warp; a thread owns rows tb + 8 i and outputs tn + 16 j, so the 16-byte reads of one
    // instruction touch 8 (x) / 4 (w) consecutive padded rows = distinct banks.  Even and odd k accumulate in the two halves of
    // a packed fp32x2 FMA (FFMA2: both operand pairs are the natural halves of the 16-byte reads) and are added at the end.
    const int tb = tid & 7, tn = tid >> 3;
    float2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int st = 0; st < nst; ++st) {
        cpa_wait<FC_ST - 2>();
        __syncthreads();
        if (st + FC_ST - 1 < nst) load(st + FC_ST - 1, (st + FC_ST - 1) % FC_ST);
        cpa_commit();
        const int buf = st % FC_ST;
        const float* xw = s_x + (buf * FC_MB + tb) * FC_LD;
        const float* ww = s_w + (buf * FC_TN + tn) * FC_LD;
#pragma unroll 4
        for (int k = 0; k < FC_KC; k += 4) {
            float4 xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xw + i * 8 * FC_LD + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(ww + j * 16 * FC_LD + k);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = __ffma2_rn(make_float2(xv[i].x, xv[i].y), make_float2(wv[j].x, wv[j].y), acc[i][j]);
                    acc[i][j] = __ffma2_rn(make_float2(xv[i].z, xv[i].w), make_float2(wv[j].z, wv[j].w), acc[i][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = tb + 8 * i;
        if (b >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn + 16 * j;
            if (n < N) atomicAdd(y + (size_t)b * N + n, acc[i][j].x + acc[i][j].y);
        }
    }
}

// ---- data gradient ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FC_NT) fc_dgrad_stream_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                                int M, int N, int K, int n_per_split) {
    extern __shared__ __align__(16) float fs[];
    float* s_w = fs;                                   // [FC_ST][FC_KC n][FC_TN + 4 k]
    float* s_d = fs + FC_ST * FC_KC * FC_LD;           // [FC_ST][FC_KC n][FC_MB + 4 b]   (dy transposed: batch contiguous)
    constexpr int DLD = FC_MB + 4;
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * FC_TN;
    const int nbeg = blockIdx.y * n_per_split, nend = min(N, nbeg + n_per_split);
    const int nst = (nend - nbeg + FC_KC - 1) / FC_KC;
    auto load = [&](int st, int buf) {
        const int nb = nbeg + st * FC_KC;
        for (int i = tid; i < FC_KC * (FC_TN / 4); i += FC_NT) {              // weights: 64 rows (n) x 16 pieces (k)
            const int r = i >> 4, c = (i & 15) * 4;
            const bool ok = nb + r < nend && k0 + c < K;
            cpa16(s_w + (buf * FC_KC + r) * FC_LD + c, w + (size_t)(ok ? nb + r : 0) * K + (ok ? k0 + c : 0), ok);
        }
    };
    // gradient tile (32 x 64 floats per stage), transposed on the way in: global -> registers before the stage's math, registers ->
    // shared memory after it, so the L2 latency of these 4-byte loads hides behind the FMAs
    constexpr int DPT = FC_MB * FC_KC / FC_NT;
    float dreg[DPT];
    auto dy_fetch = [&](int st) {
        const int nb = nbeg + st * FC_KC;
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const int i = tid + q * FC_NT, b = i / FC_KC, r = i - b * FC_KC;
            dreg[q] = (b < M && nb + r < nend) ? __ldg(dy + (size_t)b * N + nb + r) : 0.f;
        }
    };
    auto dy_store = [&](int buf) {
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const int i = tid + q * FC_NT, b = i / FC_KC, r = i - b * FC_KC;
            s_d[(buf * FC_KC + r) * DLD + b] = dreg[q];
        }
    };
    for (int s = 0; s < FC_ST - 1; ++s) {
        if (s < nst) { load(s, s); dy_fetch(s); dy_store(s); }
        cpa_commit();
    }
    const int tb = tid >> 4, tk = tid & 15;            // batch rows 4 tb .. +3, columns 4 tk .. +3
    float2 acc[4][2];                                   // [batch row][column pair]
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    for (int st = 0; st < nst; ++st) {
        cpa_wait<FC_ST - 2>();
        __syncthreads();
        const int nxt = st + FC_ST - 1;
        if (nxt < nst) { load(nxt, nxt % FC_ST); dy_fetch(nxt); }
        cpa_commit();
        const int buf = st % FC_ST;
        const float* ww = s_w + buf * FC_KC * FC_LD + tk * 4;
        const float* dd = s_d + buf * FC_KC * DLD + tb * 4;
#pragma unroll 8
        for (int r = 0; r < FC_KC; ++r) {
            const float4 wv = *reinterpret_cast<const float4*>(ww + r * FC_LD);
            const float4 dv = *reinterpret_cast<const float4*>(dd + r * DLD);
            const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
            const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 d2 = make_float2(dvv[i], dvv[i]);
                acc[i][0] = __ffma2_rn(d2, w01, acc[i][0]);
                acc[i][1] = __ffma2_rn(d2, w23, acc[i][1]);
            }
        }
        if (nxt < nst) dy_store(nxt % FC_ST);           // buffer last read in stage st-1: every thread is past this stage's barrier
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = tb * 4 + i, k = k0 + tk * 4;
        if (b < M && k < K)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dx + (size_t)b * K + k), "f"(acc[i][0].x), "f"(acc[i][0].y),
                         "f"(acc[i][1].x), "f"(acc[i][1].y)
                         : "memory");
    }
}

inline bool stream_on() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_FC_STREAM"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

}  // namespace

// return B200SP_ENOSYS when the call is outside what these kernels take (the caller then uses the split-K tensor-core GEMM)
int fc_fwd_stream(const float* x, const float* w, float* y_acc, int M, int N, int K, cudaStream_t st) {
    if (!stream_on() || M > FC_MB || K % 4 || (((uintptr_t)x | (uintptr_t)w) & 15)) return B200SP_ENOSYS;
    const size_t smem = sizeof(float) * FC_ST * (FC_TN + FC_MB) * FC_LD;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fc_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int gx = ceil_div(N, FC_TN);
    int splits = ceil_div(3 * NUM_SMS, gx);                       // ~3 CTAs per SM in flight (78 KB of shared memory each)
    const int maxs = ceil_div(K, 8 * FC_KC);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    const int kps = ceil_div(ceil_div(K, splits), FC_KC) * FC_KC;
    fc_fwd_stream_kernel<<<dim3(gx, ceil_div(K, kps)), FC_NT, smem, st>>>(x, w, y_acc, M, N, K, kps);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

int fc_dgrad_stream(const float* dy, const float* w, float* dx_acc, int M, int N, int K, cudaStream_t st) {
    if (!stream_on() || M > FC_MB || K % 4 || (((uintptr_t)dx_acc | (uintptr_t)w) & 15)) return B200SP_ENOSYS;
    const size_t smem = sizeof(float) * FC_ST * (FC_KC * FC_LD + FC_KC * (FC_MB + 4));
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fc_dgrad_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int gx = ceil_div(K, FC_TN);
    int splits = ceil_div(3 * NUM_SMS, gx);
    const int maxs = ceil_div(N, 8 * FC_KC);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    const int nps = ceil_div(ceil_div(N, splits), FC_KC) * FC_KC;
    fc_dgrad_stream_kernel<<<dim3(gx, ceil_div(N, nps)), FC_NT, smem, st>>>(dy, w, dx_acc, M, N, K, nps);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
