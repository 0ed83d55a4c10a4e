// fc_probe.cu -- standalone timing probe for the weight-streaming FC forward (fcstream.cu): which of {global access pattern, shared
// memory fill, FMA issue} bounds it.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fc_probe fc_probe.cu ; ./fc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ void cpa16(void* dst, const void* src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TN, int KC, int ST, bool MATH, bool LOADW>
__global__ void __launch_bounds__(128) fwd(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int M, int N, int K, int kps) {
    constexpr int LD = KC + 4, NT = 128, MB = 32, PW = KC / 4;
    extern __shared__ __align__(16) float fs[];
    float* s_w = fs;
    float* s_x = fs + ST * TN * LD;
    const int tid = threadIdx.x, n0 = blockIdx.x * TN;
    const int kbeg = blockIdx.y * kps, kend = min(K, kbeg + kps);
    const int nst = (kend - kbeg + KC - 1) / KC;
    auto load = [&](int st, int buf) {
        const int k0 = kbeg + st * KC;
        if (LOADW || st < ST)
            for (int i = tid; i < TN * PW; i += NT) {
                const int r = i / PW, c = (i % PW) * 4;
                const bool ok = n0 + r < N && k0 + c < kend;
                cpa16(s_w + (buf * TN + r) * LD + c, w + (size_t)(ok ? n0 + r : 0) * K + (ok ? k0 + c : 0), ok);
            }
        for (int i = tid; i < MB * PW; i += NT) {
            const int r = i / PW, c = (i % PW) * 4;
            const bool ok = r < M && k0 + c < kend;
            cpa16(s_x + (buf * MB + r) * LD + c, x + (size_t)(ok ? r : 0) * K + (ok ? k0 + c : 0), ok);
        }
    };
    for (int s = 0; s < ST - 1; ++s) { if (s < nst) load(s, s); cpa_commit(); }
    constexpr int NJ = TN / 16;
    const int tb = tid & 7, tn = tid >> 3;
    float2 acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int st = 0; st < nst; ++st) {
        cpa_wait<ST - 2>();
        __syncthreads();
        if (st + ST - 1 < nst) load(st + ST - 1, (st + ST - 1) % ST);
        cpa_commit();
        const int buf = st % ST;
        const float* xw = s_x + (buf * MB + tb) * LD;
        const float* ww = s_w + (buf * TN + tn) * LD;
        if (MATH) {
#pragma unroll 4
            for (int k = 0; k < KC; k += 4) {
                float4 xv[4], wv[NJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xw + i * 8 * LD + k);
#pragma unroll
                for (int j = 0; j < NJ; ++j) wv[j] = *reinterpret_cast<const float4*>(ww + j * 16 * LD + k);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        acc[i][j] = __ffma2_rn(make_float2(xv[i].x, xv[i].y), make_float2(wv[j].x, wv[j].y), acc[i][j]);
                        acc[i][j] = __ffma2_rn(make_float2(xv[i].z, xv[i].w), make_float2(wv[j].z, wv[j].w), acc[i][j]);
                    }
            }
        } else {
            acc[0][0].x += xw[0] + ww[0];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = tb + 8 * i;
        if (b >= M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = n0 + tn + 16 * j;
            if (n < N) atomicAdd(y + (size_t)b * N + n, acc[i][j].x + acc[i][j].y);
        }
    }
}

// direct-from-global variant: a warp owns 8 weight rows x a K range; lanes along K (coalesced 512-byte row pieces); the activations
// of the lane's 4 columns for all 32 batch rows sit in shared memory [k4][b] so that one 16-byte read gives 4 batch rows
__global__ void __launch_bounds__(256) stream_sum(const float4* __restrict__ w, float* __restrict__ y, size_t n4) {
    float4 a = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(w + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (a.x + a.y + a.z + a.w == 123.456f) y[0] = 1.f;
}

template <int TN, int KC, int ST, bool MATH, bool LOADW>
int run(const char* name, const float* x, const float* w, float* y, int M, int N, int K, int ctas_per_sm) {
    constexpr int LD = KC + 4;
    const size_t smem = sizeof(float) * ST * (TN + 32) * LD;
    auto kern = fwd<TN, KC, ST, MATH, LOADW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    const int gx = (N + TN - 1) / TN;
    int splits = (ctas_per_sm * 148 + gx - 1) / gx;
    int maxs = (K + 8 * KC - 1) / (8 * KC);
    if (splits > maxs) splits = maxs;
    const int kps = (((K + splits - 1) / splits) + KC - 1) / KC * KC;
    dim3 grid(gx, (K + kps - 1) / kps);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, 128, smem>>>(x, w, y, M, N, K, kps);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int it = 20;
    for (int i = 0; i < it; ++i) kern<<<grid, 128, smem>>>(x, w, y, M, N, K, kps);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / it;
    printf("%-34s TN=%3d KC=%3d ST=%d occ=%d grid=%dx%d smem=%zuK  %7.1f us  %6.2f TB/s\n", name, TN, KC, ST, occ, grid.x, grid.y, smem / 1024, us,
           (double)N * K * 4 / us * 1e-6);
    return 0;
}

int main() {
    const int M = 32, N = 4096, K = 9216;
    float *x, *w, *y;
    // two weight matrices alternate?  no: one 151 MB matrix exceeds the 126 MB L2, as in the real step
    CK(cudaMalloc(&x, sizeof(float) * M * K)); CK(cudaMalloc(&w, sizeof(float) * (size_t)N * K)); CK(cudaMalloc(&y, sizeof(float) * M * N));
    CK(cudaMemset(x, 0, sizeof(float) * M * K)); CK(cudaMemset(w, 0, sizeof(float) * (size_t)N * K)); CK(cudaMemset(y, 0, sizeof(float) * M * N));
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int g : {148 * 4, 148 * 8, 148 * 16}) {
            stream_sum<<<g, 256>>>((const float4*)w, y, (size_t)N * K / 4);
            cudaEventRecord(e0);
            for (int i = 0; i < 10; ++i) stream_sum<<<g, 256>>>((const float4*)w, y, (size_t)N * K / 4);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("contiguous read, grid %5d: %7.1f us  %6.2f TB/s\n", g, ms * 100, (double)N * K * 4 / (ms * 100) * 1e-6);
        }
    }
    if (run<64, 64, 3, true, true>("as shipped", x, w, y, M, N, K, 3)) return 1;
    if (run<64, 64, 3, false, true>("loads only", x, w, y, M, N, K, 3)) return 1;
    if (run<64, 64, 3, true, false>("math only (no weight loads)", x, w, y, M, N, K, 3)) return 1;
    if (run<64, 32, 4, true, true>("KC=32 4 stages", x, w, y, M, N, K, 5)) return 1;
    if (run<64, 32, 4, false, true>("KC=32 4 stages loads only", x, w, y, M, N, K, 5)) return 1;
    if (run<64, 128, 2, true, true>("KC=128 2 stages", x, w, y, M, N, K, 2)) return 1;
    if (run<64, 128, 3, false, true>("KC=128 3 stages loads only", x, w, y, M, N, K, 1)) return 1;
    if (run<128, 32, 3, true, true>("TN=128 KC=32", x, w, y, M, N, K, 3)) return 1;
    if (run<128, 32, 3, true, false>("TN=128 KC=32 math only", x, w, y, M, N, K, 3)) return 1;
    if (run<128, 64, 3, true, true>("TN=128 KC=64", x, w, y, M, N, K, 1)) return 1;
    if (run<32, 64, 4, true, true>("TN=32 KC=64", x, w, y, M, N, K, 4)) return 1;
    return 0;
}
