// mma_probe.cu -- tcgen05.mma issue-rate probe (measurement tool, not on any product path).
// One CTA per SM; one thread issues `iters` MMAs (M = 128, K = 32 bytes) back to back on whatever the shared memory holds,
// commits, waits, and reports clock64 cycles.  Variants answer "what paces one MMA on this part":
//   kind     0 tf32 | 1 bf16
//   a_src    0 A from shared memory (SS form) | 1 A from TMEM (TS form)
//   layout   0 K-major SWIZZLE_128B rows (the production layout) | 1 K-major no-swizzle (dense 8-row x 32-byte groups)
//            | 2 K-major SWIZZLE_64B | 3 K-major SWIZZLE_32B
//   rotate   0 same operand address every time | 1 walk the 4 k-steps of 4 stages like a real main loop
// Results: profiles/r2_mma_probe.txt.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int kind, int a_src, int layout, int N, int iters, int rotate, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f800000u;   // finite data
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) { tc::tmem_alloc(&tmem_slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        const uint32_t s_base = tc::smem_u32(smem);
        const uint32_t idesc = tc::make_idesc(kind == 0 ? tc::FMT_TF32 : tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K, 128, N);
        // stage = A tile (16 KB) + B tile (up to 32 KB); 3 stages
        uint32_t lt, sbo, lbo;
        if (layout == 0) { lt = 2; sbo = 1024; lbo = 0; }
        else if (layout == 1) { lt = 0; sbo = 256; lbo = 128; }
        else if (layout == 2) { lt = 4; sbo = 512; lbo = 0; }
        else { lt = 6; sbo = 256; lbo = 0; }
        const uint32_t a_tm = tmem_base + 256;             // A operand columns (TS form): beyond the accumulator
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int ks = rotate ? (i & 3) : 0, stg = rotate ? ((i >> 2) % 3) : 0;
            const uint32_t a_s = s_base + stg * 49152 + ks * 32, b_s = s_base + stg * 49152 + 16384 + ks * 32;
            const uint64_t ad = tc::smem_desc(a_s, lbo, sbo, lt), bd = tc::smem_desc(b_s, lbo, sbo, lt);
            if (a_src == 0) {
                if (kind == 0) tc::umma<true>(tmem_base, ad, bd, idesc, 1u);
                else           tc::umma<false>(tmem_base, ad, bd, idesc, 1u);
            } else {
                const uint32_t at = a_tm + (rotate ? ((i & 7) * 8) : 0);
                if (kind == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base), "r"(at), "l"(bd),
                                 "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base), "r"(at), "l"(bd),
                                 "r"(idesc), "r"(1u) : "memory");
            }
        }
        tc::umma_commit(&bar);
        uint32_t spins = 0;
        while (!tc::mbar_try_wait(&bar, 0)) { if (++spins > (1u << 26)) __trap(); }
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int b200sp_mma_probe(int kind, int a_src, int layout, int N, int iters, int rotate, int grid, long long* out_cycles, void* stream) {
    if (N < 16 || N > 256 || N % 16 || iters < 1 || grid < 1) return B200SP_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    mma_probe_kernel<<<grid, 128, 162 * 1024, (cudaStream_t)stream>>>(kind, a_src, layout, N, iters, rotate, out_cycles);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
