// tc_probe.cu -- self-test of the tcgen05 plumbing in tc_common.cuh: ONE CTA multiplies a
// 128 x Ktot by Ktot x N problem held entirely in shared memory, for every operand format the
// production kernels use (tf32 / 3xTF32 / bf16; K-major and MN-major operands; 128B swizzle).
// tests/test_tc_gpu.py compares it with float64 matmul; it pins the descriptor encodings.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

struct ProbeArgs {
    const void* A;
    const void* B;
    float* D;
    int N, nkb, mode, a_major, b_major, variant;
};

// copy one operand k-block into its swizzled row tile(s).  hi/lo: destination tiles (lo used for 3xTF32)
__device__ void probe_fill(const void* X, int MN, int Ktot, int kb, int major, int mode, uint8_t* hi, uint8_t* lo) {
    const int es = mode == 2 ? 2 : 4;
    const int KROWS = 128 / es;                 // reduction depth of one k-block
    const int natoms = major == tc::MAJOR_K ? 1 : (MN * es + 127) / 128;
    const bool b32 = major == tc::MAJOR_MN && mode != 2;
    const int rows = major == tc::MAJOR_K ? MN : KROWS;
    for (int p = threadIdx.x; p < natoms * rows * 8; p += blockDim.x) {
        const int c = p & 7, r = (p >> 3) % rows, j = (p >> 3) / rows;
        const uint8_t* src;
        if (major == tc::MAJOR_K) src = (const uint8_t*)X + ((size_t)r * Ktot) * es + kb * 128 + c * 16;
        else                      src = (const uint8_t*)X + ((size_t)(kb * KROWS + r) * MN) * es + j * 128 + c * 16;
        const uint32_t off = j * (rows * 128) + (b32 ? tc::sw128b32_off(r, c) : tc::sw128_off(r, c));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (major == tc::MAJOR_K || j * 128 + c * 16 < MN * es) v = *reinterpret_cast<const uint4*>(src);
        if (mode == 1) {
            float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
            float h[4], l[4];
            for (int i = 0; i < 4; ++i) tc::split_tf32(f[i], h[i], l[i]);
            *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            *reinterpret_cast<uint4*>(hi + off) = v;
        }
    }
}

__device__ uint64_t probe_desc(uint32_t tile, int major, int mode, int ks, int variant) {
    const int es = mode == 2 ? 2 : 4;
    const int KROWS = 128 / es;
    if (major == tc::MAJOR_K) return tc::smem_desc_sw128(tile + ks * 32, (variant & 2) ? 16 : 0, 1024);
    const uint32_t kstep_rows = 32 / es;        // rows consumed by one UMMA (8 tf32 / 16 bf16)
    uint32_t lbo = KROWS * 128, sbo = mode == 2 ? 1024 : 512;
    if (variant & 1) { uint32_t t = lbo; lbo = sbo; sbo = t; }
    return tc::smem_desc(tile + ks * kstep_rows * 128, lbo, sbo, mode == 2 ? tc::SWZ_128B : tc::SWZ_128B_BASE32B);
}

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const ProbeArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int es = g.mode == 2 ? 2 : 4;
    const int Ktot = g.nkb * (128 / es);
    const uint32_t a_bytes = 128 * 128, b_bytes = ((g.N + 31) / 32) * 32 * 128;
    // layout: per k-block  [A hi][A lo][B hi][B lo]
    const uint32_t blk = 2 * a_bytes + 2 * b_bytes;
    const int warp = threadIdx.x >> 5;
    uint32_t ncols = 32;
    while ((int)ncols < g.N) ncols <<= 1;
    if (warp == 0) { tc::tmem_alloc(&tmem_base, ncols); tc::tmem_relinquish(); }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    for (int kb = 0; kb < g.nkb; ++kb) {
        uint8_t* base = smem + kb * blk;
        probe_fill(g.A, 128, Ktot, kb, g.a_major, g.mode, base, base + a_bytes);
        probe_fill(g.B, g.N, Ktot, kb, g.b_major, g.mode, base + 2 * a_bytes, base + 2 * a_bytes + b_bytes);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tacc = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc(g.mode == 2 ? tc::FMT_BF16 : tc::FMT_TF32, g.a_major, g.b_major, 128, g.N);
        uint32_t acc = 0;
        for (int kb = 0; kb < g.nkb; ++kb) {
            const uint32_t base = tc::smem_u32(smem + kb * blk);
            const uint32_t ah = base, al = base + a_bytes, bh = base + 2 * a_bytes, bl = bh + b_bytes;
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t dah = probe_desc(ah, g.a_major, g.mode, ks, g.variant), dbh = probe_desc(bh, g.b_major, g.mode, ks, g.variant);
                if (g.mode == 2) {
                    tc::umma<false>(tacc, dah, dbh, idesc, acc);
                } else if (g.mode == 0) {
                    tc::umma<true>(tacc, dah, dbh, idesc, acc);
                } else {
                    const uint64_t dal = probe_desc(al, g.a_major, g.mode, ks, g.variant), dbl = probe_desc(bl, g.b_major, g.mode, ks, g.variant);
                    tc::umma<true>(tacc, dal, dbh, idesc, acc);
                    tc::umma<true>(tacc, dah, dbl, idesc, 1);
                    tc::umma<true>(tacc, dah, dbh, idesc, 1);
                }
                acc = 1;
            }
        }
        tc::umma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    const int row = threadIdx.x;        // TMEM lane == D row; warp w owns lanes 32w..32w+31
    for (int c0 = 0; c0 < g.N; c0 += 16) {
        uint32_t r[16];
        tc::tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        for (int i = 0; i < 16; ++i) g.D[(size_t)row * g.N + c0 + i] = __uint_as_float(r[i]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tacc, ncols);
}

}  // namespace

extern "C" int b200sp_tc_probe(const void* A, const void* B, float* D, int N, int nkb, int mode,
                               int a_major, int b_major, int variant, void* stream) {
    if (N % 16 || N < 16 || N > 256 || nkb < 1 || nkb > 2 || mode < 0 || mode > 2) return B200SP_EINVAL;
    ProbeArgs a = {A, B, D, N, nkb, mode, a_major, b_major, variant};
    const size_t smem = (size_t)nkb * (2 * 128 * 128 + 2 * ((N + 31) / 32) * 32 * 128) + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
