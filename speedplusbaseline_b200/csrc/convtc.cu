// convtc.cu -- dense k x k convolution as a "shifted GEMM" on tcgen05, operands staged by TMA.
// (Ghiasi style-transfer net, src/styleaug/ghiasi.py:6-135: 9x9 / 3x3, stride 1|2, reflection padding,
// nearest upsampling -- all folded into how the INPUT PLANES were written by b200sp_in_apply, so this
// kernel only ever sees zero-free, padding-free work.)
//
// Input: up to four bf16 "planes" [R = B*Hq*Wq pixels][C channels] (one plane for stride 1; the four
// row/column parity planes of the padded input for stride 2).  Output pixel m of the *plane grid*
// (m = (b*Hq + ph)*Wq + pw) accumulates, for every tap, plane rows m + dh*Wq + dw: a tap is a row
// SHIFT of the same 2-D matrix, so each K-chunk of the implicit GEMM is ONE TMA box
//     A_j = plane[chunk.plane][m0 + chunk.shift .. +128)[128 bytes of channels/pixels]
// landing in shared memory in the 128B-swizzled K-major UMMA format; no im2col, no SM-side operand
// work.  Narrow-channel planes put several neighbouring pixels in one 128-byte row (a tensor map with
// OVERLAPPING rows: row pitch = one pixel, row extent = 64 elements), which is how 9x9 taps over
// C=8 or C=32 planes are fed 8 or 2 taps per MMA K-chunk.  Plane-grid positions with ph >= Ho or
// pw >= Wo are padding artefacts: computed, never stored.
//
//   warp 0   : TMA producer (one elected thread), n-stage full/empty mbarrier ring
//   warp 1   : tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32 in TMEM, 128 x N tile, double-buffered)
//   warps 2-5: epilogue: tcgen05.ld -> fp32 NHWC store + per-(image, channel) sum / sum-of-squares for
//              the InstanceNorm that follows every conv of the network (warp transpose-reduce -> smem ->
//              one global atomic per channel per tile)
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int CT_NT = 192;
constexpr int CT_BM = 128;
constexpr int CT_A_BYTES = CT_BM * 128;
constexpr uint32_t CT_SMEM_LIMIT = 227 * 1024;

struct alignas(64) ConvArgs {
    CUtensorMap amap[4];
    CUtensorMap wmap;
    int32_t chunk_pc[B200SP_CONVTC_MAX_CHUNKS];      // plane | c0 << 8
    int32_t chunk_shift[B200SP_CONVTC_MAX_CHUNKS];
    int32_t n_chunks, n_stages, BN, N_out;
    int32_t R, plane_sz, Wq, Ho, Wo, num_tiles;
    int32_t OH, OW, sy, sx, oy, ox;
    uint32_t stage_bytes, off_bar, off_stat, tmem_cols;
    float* out;
    float* stats;
};

__device__ __forceinline__ void ct_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!tc::mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();          // a protocol bug traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void ct_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!tc::mbar_try_wait(bar, parity)) {
        __nanosleep(128);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void epi_bar128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// column sums of a 32 (lanes = rows) x 32 (registers = columns) tile: after the five exchange steps lane l
// holds the sum over all 32 rows of column l (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float col_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(CT_NT, 1) convtc_kernel(const __grid_constant__ ConvArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
    uint64_t* full = bars;                                  // [n_stages] TMA -> MMA
    uint64_t* empty = bars + g.n_stages;                    // [n_stages] MMA -> TMA
    uint64_t* tfull = bars + 2 * g.n_stages;                // [2] MMA -> epilogue
    uint64_t* tempty = tfull + 2;                           // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_stat = reinterpret_cast<float*>(smem + g.off_stat);   // [2 images][2][BN]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN;

    if (tid == 0) {
        for (int i = 0; i < g.n_stages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
        tc::mbar_fence_init();
        for (int i = 0; i < 4; ++i) tma::prefetch_map(&g.amap[i]);
        tma::prefetch_map(&g.wmap);
    }
    if (warp == 1) { tc::tmem_alloc(tmem_slot, g.tmem_cols); tc::tmem_relinquish(); }
    if (warp >= 2) for (int i = tid - 64; i < 4 * BN; i += 128) s_stat[i] = 0.f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // broadcast form: TMEM addresses stay in uniform registers

    if (warp == 0) {
        // ===================================== TMA PRODUCER =====================================
        if (tc::elect_one()) {       // elect.sync: the TMA / MMA issue below compiles without per-instruction waterfall loops
            int s = 0;
            uint32_t par = 1;
            const uint32_t tx = CT_A_BYTES + (uint32_t)BN * 128u;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
                const int m0 = tile * CT_BM;
                for (int j = 0; j < g.n_chunks; ++j) {
                    ct_wait(&empty[s], par);
                    const uint32_t a_s = s_base + s * g.stage_bytes, b_s = a_s + CT_A_BYTES;
                    const uint32_t bar = tc::smem_u32(&full[s]);
                    tc::mbar_arrive_expect_tx(&full[s], tx);
                    const int pc = g.chunk_pc[j];
                    tma::load_3d(a_s, &g.amap[pc & 0xff], bar, pc >> 8, 0, m0 + g.chunk_shift[j]);
                    tma::load_2d(b_s, &g.wmap, bar, j * 64, 0);
                    if (++s == g.n_stages) { s = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA ISSUER ========================================
        const uint32_t idesc = tc::make_idesc(tc::FMT_BF16, tc::MAJOR_K, tc::MAJOR_K, CT_BM, BN);
        int s = 0, ni = 0;
        uint32_t par = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++ni) {
            const int acc = ni & 1;
            ct_wait(&tempty[acc], ((ni >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int j = 0; j < g.n_chunks; ++j) {
                ct_wait(&full[s], par);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_s = s_base + s * g.stage_bytes, b_s = a_s + CT_A_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc::umma<false>(d_tmem, tc::smem_desc(a_s + ks * 32, 0, 1024, tc::SWZ_128B),
                                        tc::smem_desc(b_s + ks * 32, 0, 1024, tc::SWZ_128B), idesc, (j > 0 || ks > 0) ? 1u : 0u);
                    tc::umma_commit(&empty[s]);
                    if (j == g.n_chunks - 1) tc::umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++s == g.n_stages) { s = 0; par ^= 1; }
            }
        }
    } else {
        // ===================================== EPILOGUE ==========================================
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
        const int et = tid - 64;                          // 0..127
        const bool do_stats = g.stats != nullptr;
        int ni = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++ni) {
            const int acc = ni & 1;
            const int m0 = tile * CT_BM;
            const int m = m0 + quarter * 32 + lane;
            const int b = m / g.plane_sz;
            const int rem = m - b * g.plane_sz;
            const int ph = rem / g.Wq, pw = rem - ph * g.Wq;
            const bool valid = m < g.R && ph < g.Ho && pw < g.Wo;
            const int b_lo = m0 / g.plane_sz;
            const int m_last = min(m0 + CT_BM - 1, g.R - 1);
            const bool two = (m_last / g.plane_sz) != b_lo;
            float* orow = g.out + ((size_t)(b * g.OH + ph * g.sy + g.oy) * g.OW + pw * g.sx + g.ox) * g.N_out;
            ct_wait_sleep(&tfull[acc], (ni >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                if (BN - c0 >= 32) {
                    tc::tmem_ld32(t_row + c0, r);
                } else {
                    uint32_t r16[16];
                    tc::tmem_ld16(t_row + c0, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
                }
                tc::tmem_ld_wait();
                if (c0 + 32 >= BN) {                       // accumulator drained by this warp
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty[acc]);
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (c0 + i < g.N_out)
                            *reinterpret_cast<float4*>(orow + c0 + i) =
                                make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
                if (do_stats) {
                    for (int img = 0; img < (two ? 2 : 1); ++img) {
                        const bool sel = valid && b == b_lo + img;
                        float x[32], q[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            x[i] = sel ? __uint_as_float(r[i]) : 0.f;
                            q[i] = x[i] * x[i];
                        }
                        const float sx = col_reduce32(x, lane);
                        const float sq = col_reduce32(q, lane);
                        if (c0 + lane < BN) {
                            atomicAdd(&s_stat[(img * 2 + 0) * BN + c0 + lane], sx);
                            atomicAdd(&s_stat[(img * 2 + 1) * BN + c0 + lane], sq);
                        }
                    }
                }
            }
            if (do_stats) {
                epi_bar128();
                for (int i = et; i < 4 * BN; i += 128) {
                    const int img = i / (2 * BN), w = i - img * 2 * BN;      // w = which*BN + c
                    const float v = s_stat[i];
                    s_stat[i] = 0.f;
                    const int bi = b_lo + img;
                    if ((img == 0 || two) && bi * g.plane_sz < g.R) atomicAdd(g.stats + (size_t)bi * 2 * BN + w, v);
                }
                epi_bar128();
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 1) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

}  // namespace

extern "C" int b200sp_convtc_fwd(const b200sp_convtc_desc* d, void* stream) {
    if (!d || d->n_chunks < 1 || d->n_chunks > B200SP_CONVTC_MAX_CHUNKS) return B200SP_EINVAL;
    const int C = d->C, BN = d->N_pad;
    if (C < 8 || (C < 64 ? (64 % C) != 0 : (C % 64) != 0)) return B200SP_EINVAL;
    if (BN % 16 || BN < 16 || BN > 128 || d->N_out % 4 || d->N_out > BN) return B200SP_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(convtc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    const long long R = (long long)d->B * d->Hq * d->Wq;
    if (R + 128 + 4096 >= (1ll << 31)) return B200SP_EINVAL;
    const int cbox = C < 64 ? C : 64;
    int rc = 0;
    for (int p = 0; p < 4; ++p) {
        const void* base = d->planes[p] ? d->planes[p] : d->planes[0];
        // C >= 64: rows are pixels, 64-channel boxes.  C < 64: a 128-byte row is P = 64/C consecutive pixels, expressed
        // as OVERLAPPING rows (row pitch = one pixel < row extent): the box inner extent must equal the swizzle width
        // (a narrower inner box is padded to 128 B per inner row in shared memory).
        if (C >= 64)
            rc = tma::encode_bf16_3d(&a.amap[p], base, (uint64_t)C, 1, (uint64_t)R, (uint64_t)C * 2, (uint64_t)C * 2, 64, 1, CT_BM);
        else
            rc = tma::encode_bf16_3d(&a.amap[p], base, 64, 1, (uint64_t)R, (uint64_t)C * 2, (uint64_t)C * 2, 64, 1, CT_BM);
        if (rc) return rc;
    }
    const int Keff = 64 * d->n_chunks;
    rc = tma::encode_bf16_2d(&a.wmap, d->w, (uint64_t)Keff, (uint64_t)BN, (uint64_t)Keff * 2, 64, (uint32_t)BN);
    if (rc) return rc;
    for (int j = 0; j < d->n_chunks; ++j) {
        const b200sp_convtc_chunk& c = d->chunks[j];
        if (c.plane < 0 || c.plane > 3 || !d->planes[c.plane] || c.c0 < 0 || c.c0 + cbox > C || c.shift < 0) return B200SP_EINVAL;
        a.chunk_pc[j] = c.plane | (c.c0 << 8);
        a.chunk_shift[j] = c.shift;
    }
    a.n_chunks = d->n_chunks; a.BN = BN; a.N_out = d->N_out;
    a.R = (int)R; a.plane_sz = d->Hq * d->Wq; a.Wq = d->Wq; a.Ho = d->Ho; a.Wo = d->Wo;
    if (d->sy == 0 && d->sx == 0) { a.OH = d->Ho; a.OW = d->Wo; a.sy = a.sx = 1; a.oy = a.ox = 0; }
    else {
        a.OH = d->OH; a.OW = d->OW; a.sy = d->sy; a.sx = d->sx; a.oy = d->oy; a.ox = d->ox;
        if (a.sy < 1 || a.sx < 1 || a.oy < 0 || a.ox < 0 || (d->Ho - 1) * a.sy + a.oy >= a.OH || (d->Wo - 1) * a.sx + a.ox >= a.OW) return B200SP_EINVAL;
    }
    a.num_tiles = (int)((R + CT_BM - 1) / CT_BM);
    a.stage_bytes = CT_A_BYTES + BN * 128;
    const uint32_t fixed = 4 * BN * 4 + 512;
    int ns = (int)((CT_SMEM_LIMIT - 1024 - fixed) / a.stage_bytes);
    if (ns > 8) ns = 8;
    if (ns < 2) return B200SP_ENOSYS;
    a.n_stages = ns;
    a.off_stat = ns * a.stage_bytes;
    a.off_bar = a.off_stat + 4 * BN * 4;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * BN)) cols <<= 1;
    a.tmem_cols = cols;
    a.out = d->out; a.stats = d->stats;
    const uint32_t smem = a.off_bar + 512 + 1024;
    const int grid = a.num_tiles < NUM_SMS ? a.num_tiles : NUM_SMS;
    convtc_kernel<<<grid, CT_NT, smem, (cudaStream_t)stream>>>(a);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
