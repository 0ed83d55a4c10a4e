// tcgemm2.cu -- second-generation tcgen05 GEMM for the 1x1-convolution family (fp32 storage, 3xTF32 math).
// Same problem interface (TcgProblem), MMA issue loop and epilogue as tcgemm.cu; what changed is the OPERAND PATH,
// which bounded the first-generation kernel (475-657 SASS instructions per producer thread
// and k-block, most of them cp.async issue / address / validity arithmetic):
//
//   TMA warp (1 thread)   : cp.async.bulk.tensor boxes of the RAW fp32 operands (A [, the second tensor of a folded
//                           BatchNorm-backward operand] [, B]) into an n_raw-deep shared-memory ring; out-of-range rows /
//                           columns are zero-filled by the TMA unit, so the kernel carries no validity masks and no global
//                           address arithmetic at all.  Box layouts: K-major operand = {32 floats, rows} (dense rows of 128 B);
//                           MN-major operand = one {32 floats, 32 reduction rows} box per 128-byte atom.
//   converters (16 warps) : wait for a raw stage, ld.shared one 16-byte piece, apply the virtual-tensor transform
//                           (BN affine + activation | dy = cA*g + cB*y + cC), split fp32 -> (tf32 hi, lo) and st.shared into
//                           the 128B-swizzled UMMA operand tiles; piece <-> thread mapping is fixed, so the loop is
//                           lds -> ~30 ALU -> 2 sts per piece.
//   MMA warp / epilogue   : as in tcgemm.cu (3 tcgen05.mma kind::tf32 per k-step, correction products in their own TMEM
//                           accumulator; tcgen05.ld -> smem transpose -> coalesced global phase with the BatchNorm reductions).
//
// Small weight operands (one N tile, <= 64 KB as hi+lo tiles) are streamed through the same raw ring ONCE per CTA before
// the first A tile and stay resident.  Covers the KRN combinations: FWD (A K-major, B K-major), DGRAD (A K-major,
// B MN-major), WGRAD (A, B MN-major; split-K, fp32 red.add).  Anything else returns B200SP_ENOSYS and the caller
// falls back to tcgemm.cu.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "tcgemm.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

#ifdef TCG_TIMELINE          // debug build (B200SP_LIB_SUFFIX=_tl B200SP_NVCC_EXTRA=-DTCG_TIMELINE): clock64 stamps of CTA 0, tools/tcg2_timeline.py
__device__ long long g_tcg2_tl[14][512];
#define TL(row, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_tcg2_tl[row][idx] = clock64(); } while (0)
#else
#define TL(row, idx) do {} while (0)
#endif

namespace {

// 832 threads in two role splits (template parameter EW = epilogue warps):
//   EW =  8:  8 epilogue warps | MMA warp | TMA warp | 16 converter warps   -- main-loop-bound shapes
//   EW = 16: 16 epilogue warps | MMA warp | TMA warp |  8 converter warps   -- short-K shapes: their epilogue is a chain of dependent
//            instructions per warp (~12 cycles per instruction, schedulers 17 % busy: r3d ncu capture of dgrad [150528,24,144]),
//            so it scales with the number of warps draining chunks while the converters sit idle 70 % of the time
constexpr int MMA_T = 32, TMA_T = 32;
//   EW =  8, PW = 8 (576 threads): the register file then allows 112 registers per thread instead of 72 -- the data-gradient epilogue
//            (32 accumulator values + BatchNorm-backward operands per lane) spills and rematerialises addresses at 72
__host__ __device__ constexpr int nt_of(int ew, int pw) { return (ew + 2 + pw) * 32; }
constexpr int A_SLOTS = 2;                      // 8192-byte sub-slots of a raw 128-row x 128-byte tile
constexpr int BM = 128;
constexpr int STG_LD = 36;
constexpr int MAX_OP = 4, MAX_RAW = 8;

struct ETf { static constexpr int ES = 4, KE = 32, EPV = 4, NM = 2; static constexpr bool TF32 = true; };

struct alignas(64) Tcg2Args {
    CUtensorMap mapA, mapA2, mapB;
    b200sp_vtensor a, b;
    int P, Q, R, lda, ldb, ldo;
    int BN, numPt, numQt, splits, kb_per_split, nkb;
    int n_op, n_raw, groups;
    int ts;                          // A operand converted into TMEM and consumed by the TS form of tcgen05.mma (K-major A only)
    uint32_t a_tm_col;               // first TMEM column of the A ring (n_op slots of 64 columns: 32 hi + 32 lo)
    int stack_b;                     // 3xTF32 with [B hi ; B lo] read by one MMA of width 2*BN (2 issues per k-step instead of 3)
    int nm;                          // operand copies per tile: 2 = (tf32 hi, lo) for 3xTF32, 1 = single-pass TF32 (--use_fp16 mode)
    int b_res;
    int ac_last, nks_last;
    int a_atoms, b_atoms;            // MN-major operands: 128-byte atoms per tile that hold data (A: min(P, 128) columns)
    uint32_t a_tx, b_tx;             // TMA bytes per k-block of one A tensor / of B
    uint32_t off_bres;
    uint32_t a_op_bytes, b_op_bytes;
    uint32_t op_stage_bytes, raw_stage_bytes;
    uint32_t off_raw, off_stg, off_stat, off_bar;
    uint32_t tmem_cols;
    int nacc, acc_cols, split_epi;
    void* out;
    const float* bias;
    int out_act, has_bnf;
    b200sp_bnfwd bnf;
    const void* skip;
    float scale_out;
    int has_bnb;
    b200sp_bnbwd bnb;
    double count;
    int wait_mode;
    uint32_t epi_sleep;
};

template <int T> __device__ __forceinline__ void epi_bar_t() { asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory"); }
// Waiting on an mbarrier costs issue slots: try_wait's suspend window is short and ends on any barrier activity, so a waiting warp
// re-polls continuously (r2c ncu capture of the long-M forward: 62 % of ALL executed warp-instructions were the 16 converter
// warps polling `empty`).  Hence ONE warp per role polls and releases the others through a hardware named barrier, on
// which waiting warps issue nothing.
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// bounded waits: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity, int mode = 0) {
    uint32_t spins = 0;
    while (!tc::mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++spins > (1u << 20)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
    uint32_t spins = 0;
    while (!tc::mbar_try_wait_hint(bar, parity, 100000u)) {      // hardware-suspended; the back-off only matters if the hint expires early
        __nanosleep(ns);
        if (++spins > (1u << 22)) __trap();
    }
}

struct Item { int p0, q0, kb0, kb1; };
// Tile coordinates of the items it, it + stride, it + 2 stride, ... without divisions: it = (sp * numPt + pt) * numQt + qt is
// advanced digit by digit.  get_item's four div / mod per tile were 11 % of ALL executed warp-instructions of the long-M kernels
// (r3d ncu capture) and sit on the dependent-instruction chain of every role.
struct TileIter {
    int qt, pt, sp, dq, dp, ds;
    __device__ __forceinline__ void init(const Tcg2Args& g, int it, int stride) {
        qt = it % g.numQt;
        const int t2 = it / g.numQt;
        pt = t2 % g.numPt; sp = t2 / g.numPt;
        dq = stride % g.numQt;
        const int d2 = stride / g.numQt;
        dp = d2 % g.numPt; ds = d2 / g.numPt;
    }
    __device__ __forceinline__ void step(const Tcg2Args& g) {
        qt += dq;
        int c = 0;
        if (qt >= g.numQt) { qt -= g.numQt; c = 1; }
        pt += dp + c;
        c = 0;
        if (pt >= g.numPt) { pt -= g.numPt; c = 1; }
        sp += ds + c;
    }
    __device__ __forceinline__ Item item(const Tcg2Args& g) const {
        Item w;
        w.p0 = pt * BM;
        w.q0 = qt * g.BN;
        w.kb0 = sp * g.kb_per_split;
        w.kb1 = min(g.nkb, w.kb0 + g.kb_per_split);
        return w;
    }
};
struct KIter {
    int it, total, stride;
    TileIter ti;
    Item w;
    int kb;
    __device__ __forceinline__ void init(const Tcg2Args& g, int first, int total_, int stride_) {
        it = first; total = total_; stride = stride_;
        ti.init(g, first, stride_);
        if (it < total) { w = ti.item(g); kb = w.kb0; }
    }
    __device__ __forceinline__ bool valid() const { return it < total; }
    __device__ __forceinline__ void next(const Tcg2Args& g) {
        if (++kb >= w.kb1) {
            it += stride;
            ti.step(g);
            if (it < total) { w = ti.item(g); kb = w.kb0; }
        }
    }
};

__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- per-channel transform of a virtual tensor, mode fixed at compile time --------------------------
enum { XM_PLAIN = 0, XM_BNACT = 1, XM_DY = 2 };
struct XfP { float4 a, b, c; };
template <int MODE>
__device__ __forceinline__ void xf_load(const b200sp_vtensor& t, int ch, XfP& p) {
    if (MODE != XM_PLAIN) { p.a = ldg4(t.p0 + ch); p.b = ldg4(t.p1 + ch); }
    if (MODE == XM_DY) p.c = ldg4(t.p2 + ch);
}
template <int MODE>
__device__ __forceinline__ float4 xf_apply(float4 x, float4 x2, const XfP& p, ActP act) {
    if (MODE == XM_PLAIN) return x;
    if (MODE == XM_BNACT) {
        if (act.slope == 0.f)        // ReLU / ReLU6: kernel-uniform branch, two instructions per value after the FMA
            return make_float4(fminf(fmaxf(fmaf(x.x, p.a.x, p.b.x), 0.f), act.hi), fminf(fmaxf(fmaf(x.y, p.a.y, p.b.y), 0.f), act.hi),
                               fminf(fmaxf(fmaf(x.z, p.a.z, p.b.z), 0.f), act.hi), fminf(fmaxf(fmaf(x.w, p.a.w, p.b.w), 0.f), act.hi));
        return make_float4(act_fwd(fmaf(x.x, p.a.x, p.b.x), act), act_fwd(fmaf(x.y, p.a.y, p.b.y), act),
                           act_fwd(fmaf(x.z, p.a.z, p.b.z), act), act_fwd(fmaf(x.w, p.a.w, p.b.w), act));
    }
    return make_float4(fmaf(p.a.x, x.x, fmaf(p.b.x, x2.x, p.c.x)), fmaf(p.a.y, x.y, fmaf(p.b.y, x2.y, p.c.y)),
                       fmaf(p.a.z, x.z, fmaf(p.b.z, x2.z, p.c.z)), fmaf(p.a.w, x.w, fmaf(p.b.w, x2.w, p.c.w)));
}
// fp32 -> (tf32 hi, lo), both rounded to nearest: bit-identical to cvt.rna.tf32.f32 for finite inputs (tc_common.cuh), 5 instructions
__device__ __forceinline__ void split1(float v, float& hi, float& lo) {
    const uint32_t h = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    hi = __uint_as_float(h);
    lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
// single == true: only the rounded tf32 value is stored (one MMA per k-step instead of three)
__device__ __forceinline__ void split_store(float4 v, uint32_t dst_hi, uint32_t dst_lo, bool single) {
    if (single) {
        sts4(dst_hi, make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)));
        return;
    }
    float4 h, l;
    split1(v.x, h.x, l.x); split1(v.y, h.y, l.y); split1(v.z, h.z, l.z); split1(v.w, h.w, l.w);
    sts4(dst_hi, h);
    sts4(dst_lo, l);
}

// ---- converter: one thread's share of an operand tile ---------------------------------------------------
// The raw stage holds the TMA boxes densely: 16-byte piece q of an operand lives at byte q*16, and
//   K-major : q = row*8 + chunk                       (box {32 floats, rows})
//   MN-major: q = atom*256 + row*8 + chunk            (one {32 floats, 32 reduction rows} box per 128-byte atom)
// The operand tile uses the same (atom, row, chunk) with the UMMA swizzle applied to the chunk index.  A converter GROUP of
// GT threads (GT = 512 / groups, a multiple of 128) deals the pieces q = pg + i*GT: the chunk index and the swizzle phase of
// a thread's pieces never change (GT/8 rows is a multiple of 8), so source and destination both advance by GT*16 bytes.
template <int LAY, int MODE, bool GRP, int PROD_T>
struct Conv {
    static constexpr int APT = 1024 / PROD_T;       // 16-byte pieces of a 128-row x 128-byte tile per converter thread
    b200sp_vtensor vt;
    ActP act;
    int R, nkb, ac_last, mn_ext;
    int gc, pg, GT, npieces;
    bool single;
    uint32_t soff;
    XfP par;
    __device__ __forceinline__ void init(const b200sp_vtensor& t, int mn_ext_, int R_, int nkb_, int ac_last_, int pg_, int GT_, int tile_rows) {
        vt = t; act = act_params(t.act);
        R = R_; nkb = nkb_; ac_last = ac_last_; mn_ext = mn_ext_;
        pg = pg_; GT = GT_;
        single = false;
        gc = pg & 7;
        if (LAY == TCG_LAY_KM) {
            soff = tc::sw128_off(pg >> 3, gc);
            npieces = tile_rows * 8;
        } else {
            soff = (uint32_t)(pg >> 8) * 4096u + tc::sw128b32_off((pg >> 3) & 31, gc);
            npieces = ((tile_rows * 4 + 127) / 128) * 256;
        }
    }
    __device__ __forceinline__ void load_params(int kb) {
        if (LAY == TCG_LAY_KM && MODE != XM_PLAIN) {
            int ch = kb * 32 + gc * 4;
            ch = min(ch, R - 4);                 // columns beyond R meet zero-filled B columns: any finite value does
            xf_load<MODE>(vt, ch, par);
        }
    }
    __device__ __forceinline__ void piece(int kb, int mn0, int q, uint32_t o, uint32_t raw, uint32_t raw2, uint32_t op_hi, uint32_t op_lo) {
        const float4 r = lds4(raw + o);
        float4 r2 = f4zero();
        if (MODE == XM_DY) r2 = lds4(raw2 + o);
        float4 v;
        if (LAY == TCG_LAY_KM) {
            v = xf_apply<MODE>(r, r2, par, act);
        } else {
            v = r;
            if (MODE != XM_PLAIN) {
                int ch = mn0 + (q >> 8) * 32 + gc * 4;
                ch = min(ch, mn_ext - 4);
                xf_load<MODE>(vt, ch, par);
                v = xf_apply<MODE>(r, r2, par, act);
                // reduction rows beyond R were zero-filled by the TMA unit, but a transformed zero is not zero: mask them
                if (kb * 32 + ((q >> 3) & 31) >= R) v = f4zero();
            }
        }
        split_store(v, op_hi + soff + o, op_lo + soff + o, single);
    }
    // raw / raw2: this thread's piece 0 in the raw stage (stage base + operand offset + pg*16); op_hi / op_lo: operand tile bases
    __device__ __forceinline__ void convert(int kb, int mn0, uint32_t raw, uint32_t raw2, uint32_t op_hi, uint32_t op_lo) {
        if (LAY == TCG_LAY_KM && kb == nkb - 1 && gc >= ac_last) return;      // partial last k-block: the MMA never reads these chunks
        if (!GRP || GT == PROD_T) {         // one group: at most APT pieces per thread, fully unrolled (independent lds / sts chains)
#pragma unroll
            for (int i = 0; i < APT; ++i)
                if (pg + i * PROD_T < npieces) piece(kb, mn0, pg + i * PROD_T, (uint32_t)i * (PROD_T * 16), raw, raw2, op_hi, op_lo);
        } else if (GRP) {
            const uint32_t step = (uint32_t)GT * 16u;
            uint32_t o = 0;
#pragma unroll 2
            for (int q = pg; q < npieces; q += GT, o += step) piece(kb, mn0, q, o, raw, raw2, op_hi, op_lo);
        }
    }
};

// TMA boxes of one operand k-block into `dst` (dense, see Conv)
template <int LAY>
__device__ __forceinline__ void tma_operand(uint32_t dst, const CUtensorMap* m, uint32_t bar, int mn0, int kb, int atoms) {
    if (LAY == TCG_LAY_KM) {
        tma::load_2d(dst, m, bar, kb * 32, mn0);
    } else {
        for (int a = 0; a < atoms; ++a) tma::load_2d(dst + a * 4096, m, bar, mn0 + a * 32, kb * 32);
    }
}

// =====================================================================================================
// PRE: the operands arrive pre-split (opsplit.cu): K-major planes [hi | lo] that the TMA unit loads straight into the swizzled
// operand ring -- no raw ring, no converter warps (the kernel is launched with the first PRE_NT threads only).
template <int ALAY, int BLAY, int EPI, int AMODE, int BMODE, bool GRP, bool PRE = false, int EW = 8, int PW = 24 - EW>
__global__ void __launch_bounds__(nt_of(EW, PW), 1) tcgemm2_kernel(const __grid_constant__ Tcg2Args g) {
    constexpr int NT = nt_of(EW, PW);
    using T = float;
    using E = ETf;
    constexpr int EPI_W = EW, EPI_T = EW * 32, EPI_SETS = EW / 4, MMA_WARP = EW, TMA_WARP = EW + 1;
    constexpr int PROD_TID0 = EPI_T + MMA_T + TMA_T, PROD_T = NT - PROD_TID0;
    auto epi_bar = [] { epi_bar_t<EPI_T>(); };
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
    uint64_t* full = bars;                    // [MAX_OP]  converters -> MMA
    uint64_t* empty = bars + MAX_OP;          // [MAX_OP]  MMA -> converters
    uint64_t* tfull = bars + 2 * MAX_OP;      // [4]       MMA -> epilogue
    uint64_t* tempty = bars + 2 * MAX_OP + 4; // [4]       epilogue -> MMA
    uint64_t* bfull = bars + 2 * MAX_OP + 8;  // [1]       converters -> MMA: resident B converted
    uint64_t* rawfull = bars + 2 * MAX_OP + 9;             // [MAX_RAW] TMA -> converters
    uint64_t* rawempty = bars + 2 * MAX_OP + 9 + MAX_RAW;  // [MAX_RAW] converters -> TMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_OP + 9 + 2 * MAX_RAW);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = g.numPt * g.numQt * g.splits;

    if (tid == 0) {
        const int gw = PROD_T / 32 / (GRP ? g.groups : 1);     // warps per converter group
        for (int i = 0; i < MAX_OP; ++i) { tc::mbar_init(&full[i], PRE ? 1 : gw); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], g.split_epi ? 4 : EPI_W); }
        tc::mbar_init(bfull, PROD_T / 32);
        for (int i = 0; i < MAX_RAW; ++i) { tc::mbar_init(&rawfull[i], 1); tc::mbar_init(&rawempty[i], gw); }
        tc::mbar_fence_init();
    }
    if (warp == TMA_WARP && lane == 0) {
        tma::prefetch_map(&g.mapA);
        if (AMODE == XM_DY) tma::prefetch_map(&g.mapA2);
        tma::prefetch_map(&g.mapB);
    }
    if (warp == MMA_WARP) { tc::tmem_alloc(tmem_slot, g.tmem_cols); tc::tmem_relinquish(); }
    if (tid < EPI_T) {
        float* st = reinterpret_cast<float*>(smem + g.off_stat);
        for (int i = tid; i < EPI_W * 2 * g.BN; i += EPI_T) st[i] = 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // everything above touched only shared memory / TMEM: under programmatic dependent launch it overlapped the predecessor's tail
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // broadcast form: keeps the TMEM address (and everything derived from it) in UNIFORM registers, so tcgen05.mma needs no per-instruction R2UR waterfall
    const uint32_t b_in_stage = g.ts ? 0u : g.nm * g.a_op_bytes;
    constexpr int slotA2 = A_SLOTS, slotB = AMODE == XM_DY ? 2 * A_SLOTS : A_SLOTS;     // 8192-byte sub-slots of a raw stage: A | [A2] | B

    if (PRE && warp == TMA_WARP) {
        // ======================================= TMA PRODUCER, presplit operands ==================
        // one 3-d box per operand and k-block: {32 floats, tile rows, planes} lands as the hi tile followed by the lo tile,
        // 128-byte rows with the hardware swizzle = the K-major UMMA layout; rows / columns out of range are zero-filled
        if (tc::elect_one()) {
            int os = 0;
            uint32_t opar = 1;
            const uint32_t tx = (uint32_t)g.nm * (g.a_tx + g.b_tx);
            KIter f;
            f.init(g, blockIdx.x, total, gridDim.x);
            while (f.valid()) {
                mbar_wait_guard(&empty[os], opar);
                const uint32_t a_hi = s_base + os * g.op_stage_bytes;
                const uint32_t bar = tc::smem_u32(&full[os]);
                tc::mbar_arrive_expect_tx(&full[os], tx);
                tma::load_3d(a_hi, &g.mapA, bar, f.kb * 32, f.w.p0, 0);
                tma::load_3d(a_hi + b_in_stage, &g.mapB, bar, f.kb * 32, f.w.q0, 0);
                if (++os == g.n_op) { os = 0; opar ^= 1; }
                f.next(g);
            }
        }
    } else if (warp == TMA_WARP) {
        // ======================================= TMA PRODUCER =====================================
        if (tc::elect_one()) {
            int rs = 0;
            uint32_t par = 1;
            if (g.b_res) {
                for (int kb = 0; kb < g.nkb; ++kb) {
                    mbar_wait_guard(&rawempty[rs], par);
                    tc::mbar_arrive_expect_tx(&rawfull[rs], g.b_tx);
                    tma_operand<BLAY>(s_base + g.off_raw + rs * g.raw_stage_bytes, &g.mapB, tc::smem_u32(&rawfull[rs]), 0, kb, g.b_atoms);
                    if (++rs == g.n_raw) { rs = 0; par ^= 1; }
                }
            }
            const uint32_t tx = g.a_tx * (AMODE == XM_DY ? 2u : 1u) + (g.b_res ? 0u : g.b_tx);
            KIter f;
            f.init(g, blockIdx.x, total, gridDim.x);
            int tlt = 0;
            while (f.valid()) {
                TL(0, tlt);
                mbar_wait_guard(&rawempty[rs], par);
                TL(1, tlt);
                ++tlt;
                const uint32_t rbase = s_base + g.off_raw + rs * g.raw_stage_bytes;
                const uint32_t bar = tc::smem_u32(&rawfull[rs]);
                tc::mbar_arrive_expect_tx(&rawfull[rs], tx);
                tma_operand<ALAY>(rbase, &g.mapA, bar, f.w.p0, f.kb, g.a_atoms);
                if (AMODE == XM_DY) tma_operand<ALAY>(rbase + slotA2 * 8192, &g.mapA2, bar, f.w.p0, f.kb, g.a_atoms);
                if (!g.b_res) tma_operand<BLAY>(rbase + slotB * 8192, &g.mapB, bar, f.w.q0, f.kb, g.b_atoms);
                if (++rs == g.n_raw) { rs = 0; par ^= 1; }
                f.next(g);
            }
        }
    } else if (!PRE && warp > TMA_WARP) {
        // ======================================= CONVERTERS =======================================
        // `groups` converter groups of GT = 512/groups threads; group gi owns the k-block sequence numbers n = gi (mod groups) of
        // this CTA's stream [resident-B k-blocks | (tile, k-block) items] -- several k-blocks are in conversion at once, which is
        // what hides the lds -> ALU -> sts -> proxy fence -> arrive latency chain (~1300 cycles per k-block, r2f timelines).
        // n_raw and n_op are multiples of `groups`, so a ring slot is always filled by the same group (single-producer phases).
        const int pt = tid - PROD_TID0;
        const int G = GRP ? g.groups : 1, GT = GRP ? PROD_T / G : PROD_T;       // GRP == false: one group, everything below folds to constants
        const int gi = pt / GT, pg = pt - gi * GT;
        const bool poller = (pg >> 5) == 0;                      // first warp of the group polls the mbarriers
        Conv<ALAY, AMODE, GRP, PROD_T> CA;
        Conv<BLAY, BMODE, GRP, PROD_T> CB;
        // MN-major A (weight gradient): only the atoms that hold columns of the operand are fetched and converted -- the MMA still
        // spans 128 accumulator rows, the rest read stale shared memory and produce rows the epilogue never stores
        CA.init(g.a, g.P, g.R, g.nkb, g.ac_last, pg, GT, ALAY == TCG_LAY_KM ? BM : g.a_atoms * 32);
        CB.init(g.b, g.Q, g.R, g.nkb, g.ac_last, pg, GT, g.BN);
        CA.single = CB.single = g.nm == 1;
        const uint32_t raw0 = s_base + g.off_raw + pg * 16;
        int n = gi;                                              // sequence number of the k-block this group converts next
        const int nb = g.b_res ? g.nkb : 0;
        // ring positions advance by G per iteration and wrap at most once (every ring depth is a multiple of G >= G): no
        // divisions in the loop -- they cost as much as the conversion itself (r2l: +35 % on the weight-gradient shapes)
        int rs = gi % g.n_raw;
        uint32_t rpar = (uint32_t)(gi / g.n_raw) & 1u;
        for (; n < nb; n += G) {
            if (poller) mbar_wait_guard(&rawfull[rs], rpar);
            named_bar(2 + gi, GT);
            tc::mbar_try_wait(&rawfull[rs], rpar);              // completes at once: every thread observes the TMA phase itself
            const uint32_t b_hi = s_base + g.off_bres + n * (g.nm * g.b_op_bytes), b_lo = b_hi + g.b_op_bytes;
            CB.load_params(n);
            CB.convert(n, 0, raw0 + rs * g.raw_stage_bytes, 0, b_hi, b_lo);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&rawempty[rs]);
            rs += G;
            if (rs >= g.n_raw) { rs -= g.n_raw; rpar ^= 1u; }
        }
        if (g.b_res) {
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bfull);
        }
        KIter cons;
        cons.init(g, blockIdx.x, total, gridDim.x);
        for (int sk = nb; sk < n && cons.valid(); ++sk) cons.next(g);       // this group's first item
        int os = (n - nb) % g.n_op;
        uint32_t opar = ((uint32_t)((n - nb) / g.n_op) & 1u) ^ 1u;
        int tlc = 0;
        while (cons.valid()) {
            CA.load_params(cons.kb);
            if (!g.b_res) CB.load_params(cons.kb);
            if (pt == 0) TL(2, tlc);
            if (poller) {
                mbar_wait_guard(&rawfull[rs], rpar);
                if (pt == 0) TL(3, tlc);
                mbar_wait_guard(&empty[os], opar);
                if (pt == 0) TL(4, tlc);
            }
            if (GRP) named_bar(2 + gi, GT);
            else asm volatile("bar.sync 2, %0;" ::"n"(PROD_T) : "memory");      // immediate operands: no register-operand barrier on the default path
#ifndef TCG2_NO_RECHECK
            tc::mbar_try_wait(&rawfull[rs], rpar);              // completes at once: every thread observes the TMA phase itself
#endif
            const uint32_t rbase = raw0 + rs * g.raw_stage_bytes;
            const uint32_t a_hi = s_base + os * g.op_stage_bytes, a_lo = a_hi + g.a_op_bytes;
            const uint32_t b_hi = a_hi + b_in_stage, b_lo = b_hi + g.b_op_bytes;
            if (pt == 0) TL(11, tlc);                           // after the named barrier + phase re-check
            if (ALAY == TCG_LAY_KM && g.ts) {
                // A straight into tensor memory: warp -> (TMEM lane quarter = CTA warp index & 3, 8-column slice), lane = row.
                // The raw tile was written by TMA with the 128-byte swizzle, so 8 consecutive rows read 8 different banks groups.
                const int q = warp & 3, kq = (warp - (TMA_WARP + 1)) >> 2;
                const int kcol = cons.kb * 32 + kq * 8;
                if (kcol < g.R) {
                    const int r = q * 32 + lane;
                    const uint32_t rowb = s_base + g.off_raw + rs * g.raw_stage_bytes + r * 128;
                    XfP p0, p1;
                    const int ch = min(kcol, g.R - 8);
                    xf_load<AMODE>(g.a, ch, p0);
                    xf_load<AMODE>(g.a, ch + 4, p1);
                    const ActP act = act_params(g.a.act);
                    const uint32_t o0 = (uint32_t)(((kq * 2) ^ (r & 7)) << 4), o1 = (uint32_t)(((kq * 2 + 1) ^ (r & 7)) << 4);
                    float4 x0 = lds4(rowb + o0), x1 = lds4(rowb + o1), y0 = f4zero(), y1 = f4zero();
                    if (AMODE == XM_DY) { y0 = lds4(rowb + slotA2 * 8192 + o0); y1 = lds4(rowb + slotA2 * 8192 + o1); }
                    const float4 v0 = xf_apply<AMODE>(x0, y0, p0, act), v1 = xf_apply<AMODE>(x1, y1, p1, act);
                    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float hi, lo;
                        split1(vv[i], hi, lo);
                        h[i] = __float_as_uint(hi); l[i] = __float_as_uint(lo);
                    }
                    const uint32_t ta = tmem_base + g.a_tm_col + os * 64 + kq * 8 + ((uint32_t)(q * 32) << 16);
                    tc::tmem_st8(ta, h);
                    if (g.nm == 2) tc::tmem_st8(ta + 32, l);
                }
                tc::tmem_st_wait();
                tc::tc_fence_before();
            } else {
                CA.convert(cons.kb, cons.w.p0, rbase, rbase + slotA2 * 8192, a_hi, a_lo);
            }
            if (!g.b_res) CB.convert(cons.kb, cons.w.q0, rbase + slotB * 8192, 0, b_hi, b_lo);
            if (pt == 0) TL(12, tlc);                           // conversions issued
            tc::fence_proxy_async_smem();
            if (pt == 0) TL(13, tlc);                           // proxy fence done
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&full[os]); tc::mbar_arrive(&rawempty[rs]); }
            if (pt == 0) TL(5, tlc);
            ++tlc;
            rs += G;
            if (rs >= g.n_raw) { rs -= g.n_raw; rpar ^= 1u; }
            os += G;
            if (os >= g.n_op) { os -= g.n_op; opar ^= 1u; }
            if (G == 1) cons.next(g);
            else for (int sk = 0; sk < G && cons.valid(); ++sk) cons.next(g);
        }
    } else if (warp == MMA_WARP) {
        // ======================================= MMA ISSUER ======================================
        const uint32_t idesc = tc::make_idesc(E::TF32 ? tc::FMT_TF32 : tc::FMT_BF16, ALAY == TCG_LAY_MM, BLAY == TCG_LAY_MM, BM, g.BN);
        const uint32_t idesc2 = tc::make_idesc(tc::FMT_TF32, ALAY == TCG_LAY_MM, BLAY == TCG_LAY_MM, BM, 2 * g.BN);
        // per-k-step start-address advance (bytes) and LBO / SBO / layout of each operand
        constexpr uint32_t KSTEP_KM = 32, KSTEP_MM = (E::TF32 ? 8 : 16) * 128;
        constexpr uint32_t LBO_MM = E::KE * 128, SBO_MM = E::TF32 ? 512 : 1024;
        constexpr uint32_t LT_MM = E::TF32 ? tc::SWZ_128B_BASE32B : tc::SWZ_128B;
        // shared-memory descriptors: the high word is a per-operand constant, the low word is (address >> 4) plus the
        // leading-byte-offset field -- per k-step the issuing thread only adds a constant (it is the serial bottleneck
        // of short tiles, so every instruction here counts)
        constexpr uint32_t A_STEP = (ALAY == TCG_LAY_KM ? KSTEP_KM : KSTEP_MM) >> 4, B_STEP = (BLAY == TCG_LAY_KM ? KSTEP_KM : KSTEP_MM) >> 4;
        constexpr uint32_t A_LBO = ALAY == TCG_LAY_KM ? 0u : ((LBO_MM >> 4) << 16), B_LBO = BLAY == TCG_LAY_KM ? 0u : ((LBO_MM >> 4) << 16);
        constexpr uint32_t A_HIW = ((ALAY == TCG_LAY_KM ? 1024u : SBO_MM) >> 4) | (1u << 14) | ((uint32_t)(ALAY == TCG_LAY_KM ? tc::SWZ_128B : LT_MM) << 29);
        constexpr uint32_t B_HIW = ((BLAY == TCG_LAY_KM ? 1024u : SBO_MM) >> 4) | (1u << 14) | ((uint32_t)(BLAY == TCG_LAY_KM ? tc::SWZ_128B : LT_MM) << 29);
        auto mk = [](uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); };
        if (g.b_res) { mbar_wait_guard(bfull, 0, g.wait_mode); tc::tc_fence_after(); }
        int os = 0, acc = 0, tlm = 0;
        uint32_t fpar = 0, tpar = 1;
        TileIter ti;
        ti.init(g, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < total; it += gridDim.x, ti.step(g)) {
            const Item w = ti.item(g);
            mbar_wait_guard(&tempty[acc], tpar, g.wait_mode);
            tc::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * g.acc_cols;
            // 3xTF32: the two correction products (lo*hi, hi*lo) go to a SECOND accumulator and are added in fp32-RN by the
            // epilogue.  tcgen05 accumulates with round-toward-zero, so every MMA into the large accumulator costs up to one
            // ulp of bias; keeping the small terms out of it cuts the number of such truncations from 3K/8 to K/8.
            const uint32_t d_corr = d_tmem + (g.acc_cols > g.BN ? g.BN : 0);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                if (lane == 0) TL(6, tlm);
                mbar_wait_guard(&full[os], fpar, g.wait_mode);
                tc::tc_fence_after();
                if (lane == 0) TL(7, tlm);
                if (tc::elect_one()) {
                    const uint32_t a_hi = s_base + os * g.op_stage_bytes, a_lo = a_hi + g.a_op_bytes;
                    const uint32_t b_hi = g.b_res ? s_base + g.off_bres + kb * (g.nm * g.b_op_bytes) : a_hi + b_in_stage;
                    const uint32_t b_lo = b_hi + g.b_op_bytes;
                    // a K-major operand only holds the active chunks of a partial last k-block: never read beyond them
                    const int nks = ((ALAY == TCG_LAY_KM || BLAY == TCG_LAY_KM) && kb == g.nkb - 1) ? g.nks_last : 4;
                    const uint32_t al = (a_lo >> 4) + A_LBO, ah = (a_hi >> 4) + A_LBO, bl = (b_lo >> 4) + B_LBO, bh = (b_hi >> 4) + B_LBO;
                    const uint32_t first = kb > w.kb0 ? 1u : 0u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (ks < nks) {
                            const uint32_t accum = ks > 0 ? 1u : first;
                            if (ALAY == TCG_LAY_KM && g.ts) {
                                const uint32_t at = tmem_base + g.a_tm_col + os * 64 + ks * 8;
                                if (g.stack_b) {
                                    tc::umma_ts_tf32(d_tmem, at, mk(bh + ks * B_STEP, B_HIW), idesc2, accum);
                                    tc::umma_ts_tf32(d_corr, at + 32, mk(bh + ks * B_STEP, B_HIW), idesc, 1u);
                                } else if (g.nm == 2) {
                                    tc::umma_ts_tf32(d_corr, at + 32, mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                                    tc::umma_ts_tf32(d_corr, at, mk(bl + ks * B_STEP, B_HIW), idesc, 1u);
                                    tc::umma_ts_tf32(d_tmem, at, mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                                } else {
                                    tc::umma_ts_tf32(d_tmem, at, mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                                }
                            } else if (g.stack_b) {
                                // [B hi ; B lo] are adjacent tiles: ONE MMA of width 2*BN computes A_hi*B_hi into the main accumulator
                                // columns and A_hi*B_lo into the correction columns right behind them; a second one adds A_lo*B_hi
                                // to the correction columns -- 2 issues per k-step instead of 3 for the same tensor work (the issuing
                                // thread, ~55 cycles per UTCHMMA, is on the critical path of every k-block: DESIGN.md 3.10)
                                tc::umma<true>(d_tmem, mk(ah + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc2, accum);
                                tc::umma<true>(d_corr, mk(al + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, 1u);
                            } else if (g.nm == 2) {
                                const bool split = g.acc_cols > g.BN;
                                tc::umma<true>(d_corr, mk(al + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                                tc::umma<true>(d_corr, mk(ah + ks * A_STEP, A_HIW), mk(bl + ks * B_STEP, B_HIW), idesc, 1u);
                                tc::umma<true>(d_tmem, mk(ah + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, split ? accum : 1u);
                            } else {        // single-pass TF32: one MMA per k-step on the rounded operands
                                tc::umma<true>(d_tmem, mk(ah + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                            }
                        }
                    }
                    tc::umma_commit(&empty[os]);
                    if (kb == w.kb1 - 1) tc::umma_commit(&tfull[acc]);
                    TL(8, tlm);
                }
                ++tlm;
                __syncwarp();
                if (++os == g.n_op) { os = 0; fpar ^= 1; }
            }
            if (++acc == g.nacc) { acc = 0; tpar ^= 1; }
        }
    } else {
        // ======================================= EPILOGUE ========================================
        const int lq = warp & 3, half = warp >> 2;      // TMEM lane quarter, warp set (0 .. EPI_SETS-1)
        float* stg = reinterpret_cast<float*>(smem + g.off_stg) + warp * (32 * STG_LD);
        const uint32_t stg_u = tc::smem_u32(stg);
        float* stat_all = reinterpret_cast<float*>(smem + g.off_stat);      // [EPI_W][2][BN]
        float* stat = stat_all + warp * (2 * g.BN);
        const bool do_stats = (EPI == TCG_EPI_FWD && g.has_bnf) || (EPI == TCG_EPI_DGRAD && g.has_bnb && g.bnb.s1 != nullptr);
        const bool plain_out = EPI == TCG_EPI_FWD && g.bias == nullptr && g.out_act == B200SP_ACT_NONE;
        const int cq = lane & 7, rs = lane >> 3;
        const ActP oact = act_params(EPI == TCG_EPI_FWD ? g.out_act : g.bnb.act);
        T* outT = reinterpret_cast<T*>(g.out);
        float* outF = reinterpret_cast<float*>(g.out);
        const int nchunks = (g.BN + 31) >> 5;
        // The two warp sets (4 TMEM lane quarters each) share a tile by 32-column chunks, chunk ci to set (ci + tile index) & 1:
        // with an odd chunk count (96- and 80-wide tiles of the long-M layers) a fixed assignment left one set with twice the
        // work of the other on EVERY tile (r2f timeline: 5.4 k cycles per 128x96 tile, 2 chunks on the critical path); rotated,
        // each set drains 3 chunks per 2 tiles.  The sets wait for their accumulator independently, so one can run a tile ahead.
        const int c_step = g.split_epi ? 1 : EPI_SETS;
        int ni = 0;
        int cur_q0 = -1;
        auto flush = [&](int q0) {
            epi_bar();
            for (int c = tid; c < g.BN; c += EPI_T) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int w = 0; w < EPI_W; ++w) {
                    a += (double)stat_all[w * 2 * g.BN + c];
                    b += (double)stat_all[w * 2 * g.BN + g.BN + c];
                    stat_all[w * 2 * g.BN + c] = 0.f;
                    stat_all[w * 2 * g.BN + g.BN + c] = 0.f;
                }
                if (q0 + c < g.Q) {
                    if (EPI == TCG_EPI_FWD) { atomicAdd(g.bnf.sum + q0 + c, a); atomicAdd(g.bnf.sumsq + q0 + c, b); }
                    else                    { atomicAdd(g.bnb.s1 + q0 + c, a);  atomicAdd(g.bnb.s2 + q0 + c, b); }
                }
            }
            epi_bar();
        };
        int acc = -1;
        int tle = 0;
        uint32_t tpar = 1;
        TileIter ti;
        ti.init(g, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < total; it += gridDim.x, ++ni, ti.step(g)) {
            if (++acc == g.nacc) acc = 0;
            if (acc == 0) tpar ^= 1;
            if (g.split_epi && (ni & (EPI_SETS - 1)) != half) continue;      // another warp set drains this tile
            const Item w = ti.item(g);
            if (do_stats && cur_q0 >= 0 && cur_q0 != w.q0) flush(cur_q0);
            cur_q0 = w.q0;
            const int c_first = g.split_epi ? 0 : ((half + ni) & (EPI_SETS - 1));
            const int last_chunk = c_first < nchunks ? c_first + ((nchunks - 1 - c_first) & ~(c_step - 1)) : -1;      // c_step is a power of two
            if (EPI == TCG_EPI_DGRAD && (g.has_bnb || g.skip)) {
                // the saved conv output y (activation mask / BN reductions) and the skip gradient of this tile are pulled into L2
                // now, while the main loop still runs: the epilogue's loads then cost an L2 hit instead of a DRAM round trip
                // (r2g timeline: 7-9 k cycles per column chunk were spent waiting on them)
                for (int ci = c_first; ci < nchunks; ci += c_step) {
                    const int col = w.q0 + ci * 32 + 4 * cq;
                    if (col < g.Q) {
#pragma unroll
                        for (int ps = 0; ps < 8; ++ps) {
                            const int row = w.p0 + lq * 32 + rs + ps * 4;
                            if (row < g.P) {
                                const size_t off = (size_t)row * g.ldo + col;
                                if (g.has_bnb) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const T*>(g.bnb.y) + off));
                                if (g.skip) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const T*>(g.skip) + off));
                            }
                        }
                    }
                }
            }
            {   // one polling warp per epilogue set, the others wait on a named barrier (ids 6 / 7; 2..5 belong to the converter groups)
                if ((warp & 3) == 0) mbar_wait_sleep(&tfull[acc], tpar, g.epi_sleep);
                named_bar(6 + half, 128);
            }
            if (tid == 0) TL(9, 2 * ni);
            tc::tc_fence_after();
            const uint32_t t_row = tmem_base + acc * g.acc_cols + ((uint32_t)(lq * 32) << 16);
            const bool split_acc = g.acc_cols > g.BN;
            if (last_chunk < 0) {                        // nothing to read for this warp: release immediately
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            }
            for (int ci = c_first; ci < nchunks; ci += c_step) {
                const int c0 = ci * 32;
                const int ncol = min(32, g.BN - c0);
                uint32_t r[32];
                if (tid == 0) { TL(10, tle); ++tle; }
                // the main accumulator and the first half of the correction accumulator are requested together (one wait for both)
                uint32_t q16[16];
                if (ncol == 32) {
                    tc::tmem_ld32(t_row + c0, r);
                } else {
                    uint32_t r16[16];
                    tc::tmem_ld16(t_row + c0, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
                }
                constexpr bool MERGE_LD = EPI != TCG_EPI_DGRAD || NT <= 576;      // the dgrad epilogue is register-bound at 72 registers: it keeps the loads apart
                if (MERGE_LD && split_acc) tc::tmem_ld16(t_row + g.BN + c0, q16);
                tc::tmem_ld_wait();
                if (tid == 0) { TL(10, tle); ++tle; }
                if (split_acc) {                // add the correction accumulator (fp32 round-to-nearest), 16 columns at a time
                    if (!MERGE_LD) { tc::tmem_ld16(t_row + g.BN + c0, q16); tc::tmem_ld_wait(); }
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(q16[i]));
                    if (16 < ncol) {
                        tc::tmem_ld16(t_row + g.BN + c0 + 16, q16);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[16 + i] = __float_as_uint(__uint_as_float(r[16 + i]) + __uint_as_float(q16[i]));
                    }
                }
                if (tid == 0) { TL(10, tle); ++tle; }
                if (ci == last_chunk) {         // accumulator drained by this warp: hand TMEM back to the MMA warp
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty[acc]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts4(stg_u + (lane * STG_LD + 4 * j) * 4,
                         make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
                __syncwarp();
                if (tid == 0) { TL(10, tle); ++tle; }
                // ---- coalesced phase: lane = (row sub-index rs, column quad cq) ----
                const int col = w.q0 + c0 + 4 * cq;
                const bool cok = 4 * cq < ncol && col < g.Q;
                float4 bias4 = f4zero(), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = f4zero(), mu4 = f4zero(), rs4 = f4zero();
                if (cok) {
                    if (EPI == TCG_EPI_FWD && g.bias) bias4 = ldg4(g.bias + col);
                    if (EPI == TCG_EPI_DGRAD && g.has_bnb) {
                        if (g.bnb.scale) { sc4 = ldg4(g.bnb.scale + col); sh4 = ldg4(g.bnb.shift + col); }
                        if (g.bnb.s1) { mu4 = ldg4(g.bnb.mean + col); rs4 = ldg4(g.bnb.rstd + col); }
                    }
                }
                float4 ls = f4zero(), lq4 = f4zero();
                const int row_base = w.p0 + lq * 32 + rs;
                // rows are handled in two batches of four so that the global loads of the dgrad epilogue (saved conv
                // output y for the activation mask / BN reductions, skip gradient) are all in flight together
                // instead of one dependent round trip per row
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float4 yv[4], sv[4];
                    if (EPI == TCG_EPI_DGRAD) {
#pragma unroll
                        for (int p4 = 0; p4 < 4; ++p4) {
                            const int row = row_base + (hb * 4 + p4) * 4;
                            const bool ok = cok && row < g.P;
                            const size_t off = (size_t)row * g.ldo + col;
                            yv[p4] = (g.has_bnb && ok) ? Vec4<T>::ld(reinterpret_cast<const T*>(g.bnb.y) + off) : f4zero();
                            sv[p4] = (g.skip && ok) ? Vec4<T>::ld(reinterpret_cast<const T*>(g.skip) + off) : f4zero();
                        }
                    }
#pragma unroll
                    for (int p4 = 0; p4 < 4; ++p4) {
                        const int ps = hb * 4 + p4;
                        const int trow = ps * 4 + rs;
                        const int row = row_base + ps * 4;
                        if (!(cok && row < g.P)) continue;
                        float4 v = lds4(stg_u + (trow * STG_LD + 4 * cq) * 4);
                        const size_t off = (size_t)row * g.ldo + col;
                        if (EPI == TCG_EPI_FWD) {
                            if (!plain_out) {
                                v.x = act_fwd(v.x + bias4.x, oact); v.y = act_fwd(v.y + bias4.y, oact);
                                v.z = act_fwd(v.z + bias4.z, oact); v.w = act_fwd(v.w + bias4.w, oact);
                            }
                            Vec4<T>::st(outT + off, v);
                            if (!E::TF32) {      // statistics of the value as stored (bf16-rounded)
                                v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
                                v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
                            }
                            ls.x += v.x; ls.y += v.y; ls.z += v.z; ls.w += v.w;
                            lq4.x = fmaf(v.x, v.x, lq4.x); lq4.y = fmaf(v.y, v.y, lq4.y); lq4.z = fmaf(v.z, v.z, lq4.z); lq4.w = fmaf(v.w, v.w, lq4.w);
                        } else if (EPI == TCG_EPI_DGRAD) {
                            v.x *= g.scale_out; v.y *= g.scale_out; v.z *= g.scale_out; v.w *= g.scale_out;
                            const float4 sk = sv[p4];
                            v.x += sk.x; v.y += sk.y; v.z += sk.z; v.w += sk.w;
                            if (g.has_bnb) {
                                const float4 y = yv[p4];
                                v.x *= act_bwd(fmaf(y.x, sc4.x, sh4.x), oact); v.y *= act_bwd(fmaf(y.y, sc4.y, sh4.y), oact);
                                v.z *= act_bwd(fmaf(y.z, sc4.z, sh4.z), oact); v.w *= act_bwd(fmaf(y.w, sc4.w, sh4.w), oact);
                                ls.x += v.x; ls.y += v.y; ls.z += v.z; ls.w += v.w;
                                lq4.x = fmaf(v.x, (y.x - mu4.x) * rs4.x, lq4.x); lq4.y = fmaf(v.y, (y.y - mu4.y) * rs4.y, lq4.y);
                                lq4.z = fmaf(v.z, (y.z - mu4.z) * rs4.z, lq4.z); lq4.w = fmaf(v.w, (y.w - mu4.w) * rs4.w, lq4.w);
                            }
                            Vec4<T>::st(outT + off, v);
                        } else {
                            // one 16-byte vector reduction instead of four scalar atomics (sm_90+: red.global.add.v4.f32)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(outF + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                         : "memory");
                        }
                    }
                }
                if (tid == 0) { TL(10, tle); ++tle; }
                if (do_stats) {
#pragma unroll
                    for (int o = 8; o < 32; o <<= 1) {
                        ls.x += __shfl_xor_sync(0xffffffffu, ls.x, o); ls.y += __shfl_xor_sync(0xffffffffu, ls.y, o);
                        ls.z += __shfl_xor_sync(0xffffffffu, ls.z, o); ls.w += __shfl_xor_sync(0xffffffffu, ls.w, o);
                        lq4.x += __shfl_xor_sync(0xffffffffu, lq4.x, o); lq4.y += __shfl_xor_sync(0xffffffffu, lq4.y, o);
                        lq4.z += __shfl_xor_sync(0xffffffffu, lq4.z, o); lq4.w += __shfl_xor_sync(0xffffffffu, lq4.w, o);
                    }
                    if (rs == 0 && 4 * cq < ncol) {
                        float* s0 = stat + c0 + 4 * cq;
                        float* s1 = stat + g.BN + c0 + 4 * cq;
                        s0[0] += ls.x; s0[1] += ls.y; s0[2] += ls.z; s0[3] += ls.w;
                        s1[0] += lq4.x; s1[1] += lq4.y; s1[2] += lq4.z; s1[3] += lq4.w;
                    }
                }
                __syncwarp();      // staging tile is reused by the next column chunk
            }
            if (tid == 0) TL(9, 2 * ni + 1);
        }
        if (do_stats) {
            flush(cur_q0 >= 0 ? cur_q0 : 0);       // every epilogue warp takes part (bar.sync), even one that drained no tile
            // elect the last CTA of the grid: it turns the accumulated sums into scale/shift (fwd) or dy coefficients (bwd)
            __threadfence();
            epi_bar();
            if (tid == 0) {
                uint32_t* ticket = EPI == TCG_EPI_FWD ? g.bnf.ticket : g.bnb.ticket;
                const uint32_t t = atomicAdd(ticket, 1u);
                const int last = (t == gridDim.x - 1);
                if (last) *ticket = 0u;
                *s_flag = last;
            }
            epi_bar();
            if (*s_flag) {
                __threadfence();
                if (EPI == TCG_EPI_FWD) bn_fwd_finalize_all(g.bnf, g.Q, g.count, tid, EPI_T);
                else                    bn_bwd_finalize_all(g.bnb, g.Q, g.count, tid, EPI_T);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == MMA_WARP) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// -------------------------------------------------------------------------------------------------
constexpr uint32_t SMEM_LIMIT = 227 * 1024;
constexpr uint32_t BRES_LIMIT = 64 * 1024;

// raw fp32 operand stored [rows][ld] with `cols` valid columns: box {32 floats, box_rows}, no swizzle, zero fill
inline int encode_f32(CUtensorMap* m, const void* base, int cols, int rows, int ld, int box_rows, bool swz128 = false) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return B200SP_ENOSYS;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : B200SP_ENOSYS;
}

// dry: only plan (tile width, resident B, ring depths -> a), no tensor maps, no launch
template <int ALAY, int BLAY, int EPI, int AMODE, int BMODE, int EW = 8, int PW = 24 - EW>
int launch_cfg(Tcg2Args& a, cudaStream_t st, bool dry = false) {
    constexpr int EPI_W = EW, EPI_SETS = EW / 4, NT = nt_of(EW, PW);
    // converter groups (several k-blocks in conversion at once) are an opt-in experiment: B200SP_TCG2_GROUPS=2|4.  Measured
    // (profiles/r2_gemm_variants.txt): +2-4 % on the long-M data-gradient shapes, nothing elsewhere, while the run-time group
    // geometry costs the single-group path 2.4x more converter instructions -- so the default instantiation has it compiled out.
    using E = ETf;
    static bool attr_set = false;
    auto kern = tcgemm2_kernel<ALAY, BLAY, EPI, AMODE, BMODE, false, false, EW, PW>;
    // converter groups exist for the 16-converter-warp split only
    auto kern_g = tcgemm2_kernel<ALAY, BLAY, EPI, AMODE, BMODE, EW == 8 && PW == 16, false, EW, PW>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int numPt = ceil_div(a.P, BM);
    a.nkb = ceil_div(a.R, E::KE);
    a.a_atoms = a.P >= BM ? BM / 32 : ceil_div(a.P, 32);
    {
        const int rem = a.R - (a.nkb - 1) * E::KE;
        a.nks_last = ceil_div(rem, 2 * E::EPV);
        a.ac_last = 2 * a.nks_last;
    }
    {
        static int ts_env = -1;
        if (ts_env < 0) { const char* e = getenv("B200SP_TCG2_TS"); ts_env = (e && e[0] == '1') ? 1 : 0; }
        a.ts = (ts_env && EW == 8 && PW == 16 && ALAY == TCG_LAY_KM && a.R >= 8 && a.R % 8 == 0) ? 1 : 0;
    }
    // ---- tile width: the widest that fits shared memory; narrower while the grid does not cover the machine ----
    bool fits = false;
    for (int cap = 128; cap >= 32 && !fits; cap -= (cap > 64 ? 32 : 16)) {
        int numQt = ceil_div(a.Q, cap);
        int BN = ceil_div(ceil_div(a.Q, numQt), 16) * 16;
        if (EPI != TCG_EPI_ATOMIC) {
            while (numPt * numQt < NUM_SMS && BN > 32) {
                ++numQt;
                BN = ceil_div(ceil_div(a.Q, numQt), 16) * 16;
            }
        }
        numQt = ceil_div(a.Q, BN);
        a.BN = BN; a.numPt = numPt; a.numQt = numQt;
        a.a_op_bytes = BM * 128;
        a.b_atoms = ceil_div(BN * E::ES, 128);
        a.b_op_bytes = BLAY == TCG_LAY_KM ? BN * 128 : a.b_atoms * E::KE * 128;
        const int nb_slots = ceil_div((int)a.b_op_bytes, 8192);
        const uint32_t bres_bytes = (uint32_t)a.nkb * a.nm * a.b_op_bytes;
        a.b_res = (EPI != TCG_EPI_ATOMIC && numQt == 1 && BMODE == XM_PLAIN && bres_bytes <= BRES_LIMIT) ? 1 : 0;
        a.op_stage_bytes = a.nm * ((a.ts ? 0 : a.a_op_bytes) + (a.b_res ? 0 : a.b_op_bytes));
        a.raw_stage_bytes = (A_SLOTS + (AMODE == XM_DY ? A_SLOTS : 0) + (a.b_res ? 0 : nb_slots)) * 8192;
        if (a.b_res && a.raw_stage_bytes < (uint32_t)nb_slots * 8192) a.raw_stage_bytes = nb_slots * 8192;
        const uint32_t fixed = EPI_W * 32 * STG_LD * 4 + EPI_W * 2 * BN * 4 + 512 + (a.b_res ? bres_bytes : 0);
        auto fit = [&](int nop, int nraw) { return nop * a.op_stage_bytes + nraw * a.raw_stage_bytes + fixed + 1088 <= SMEM_LIMIT; };
        if (!fit(2, 2)) continue;
        fits = true;
        // converter groups: as many as the rings allow (every ring depth is a multiple of the group count), then deeper rings:
        // raw stages first (they hide the TMA latency and are the cheaper ones), operand stages after
        static int gmax = -1;
        if (gmax < 0) { const char* e = getenv("B200SP_TCG2_GROUPS"); gmax = e ? atoi(e) : 1; if (gmax != 1 && gmax != 2 && gmax != 4) gmax = 1; }
        static int gmax_wg = -1;      // weight gradient: its own switch (B200SP_TCG2_WG_GROUPS)
        if (gmax_wg < 0) { const char* e = getenv("B200SP_TCG2_WG_GROUPS"); gmax_wg = e ? atoi(e) : 1; if (gmax_wg != 1 && gmax_wg != 2 && gmax_wg != 4) gmax_wg = 1; }
        int G = (a.ts || EW != 8 || PW != 16) ? 1 : (EPI == TCG_EPI_ATOMIC ? gmax_wg : gmax);        // the TMEM A path deals one k-block to all 16 converter warps
        while (G > 1 && !fit(G, G)) G >>= 1;
        if (G == 1) { a.n_op = 2; a.n_raw = 2; } else { a.n_op = G; a.n_raw = G; }
        a.groups = G;
        const int step = G;
        while (a.n_raw + step <= MAX_RAW && a.n_raw < 2 * a.n_op && fit(a.n_op, a.n_raw + step)) a.n_raw += step;
        while (a.n_op + step <= MAX_OP && fit(a.n_op + step, a.n_raw)) a.n_op += step;
        while (a.n_raw + step <= MAX_RAW && fit(a.n_op, a.n_raw + step)) a.n_raw += step;
        a.off_bres = a.n_op * a.op_stage_bytes;
        a.off_raw = a.off_bres + (a.b_res ? bres_bytes : 0);
        a.off_stg = a.off_raw + a.n_raw * a.raw_stage_bytes;
        a.off_stat = a.off_stg + EPI_W * 32 * STG_LD * 4;
        a.off_bar = (a.off_stat + EPI_W * 2 * BN * 4 + 15) & ~15u;
    }
    if (!fits) return B200SP_ENOSYS;
    const int numQt = a.numQt, BN = a.BN;
    a.splits = 1;
    if (EPI == TCG_EPI_ATOMIC) {
        const int base = numPt * numQt;
        int s = ceil_div(2 * NUM_SMS, base);
        const int smax = a.nkb / 4 > 0 ? a.nkb / 4 : 1;
        if (s > smax) s = smax;
        if (s < 1) s = 1;
        a.splits = s;
    }
    a.kb_per_split = ceil_div(a.nkb, a.splits);
    a.splits = ceil_div(a.nkb, a.kb_per_split);
    const uint32_t smem = a.off_bar + 512 + 1024;
    a.acc_cols = (a.nm == 2 && 2 * 2 * BN <= 512) ? 2 * BN : BN;
    a.nacc = 4 * a.acc_cols <= 512 ? 4 : 2;
    if (a.ts) {                      // tensor memory holds the A ring too: n_op slots of 64 columns behind the accumulators
        int nacc = (512 - a.n_op * 64) / a.acc_cols;
        if (nacc < 1) return B200SP_ENOSYS;
        a.nacc = nacc > 4 ? 4 : nacc;
        a.a_tm_col = (uint32_t)(a.nacc * a.acc_cols);
    }
    {
        static int sb = -1;
        if (sb < 0) { const char* e = getenv("B200SP_TCG2_STACKB"); sb = (e && e[0] == '0') ? 0 : 1; }
        // MN-major B: the hi tile must end on an atom boundary for the lo tile to continue the N index
        a.stack_b = (sb && a.nm == 2 && a.acc_cols == 2 * BN && 2 * BN <= 256 && (BLAY == TCG_LAY_KM || BN % 32 == 0)) ? 1 : 0;
    }
    a.split_epi = (BN <= 32 && numQt == 1 && EPI_SETS >= 2) ? 1 : 0;
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.nacc * a.acc_cols)) cols <<= 1;
    a.tmem_cols = a.ts ? 512u : cols;
    if (dry) return 0;
    // ---- tensor maps of the raw operands ----
    int rc;
    if (ALAY == TCG_LAY_KM) { rc = encode_f32(&a.mapA, a.a.x, a.R, a.P, a.lda, BM, a.ts != 0); a.a_tx = BM * 128; }
    else                    { rc = encode_f32(&a.mapA, a.a.x, a.P, a.R, a.lda, 32); a.a_tx = a.a_atoms * 4096; }
    if (rc) return rc;
    if (AMODE == XM_DY) {
        rc = ALAY == TCG_LAY_KM ? encode_f32(&a.mapA2, a.a.x2, a.R, a.P, a.lda, BM, a.ts != 0) : encode_f32(&a.mapA2, a.a.x2, a.P, a.R, a.lda, 32);
        if (rc) return rc;
    } else {
        a.mapA2 = a.mapA;
    }
    if (BLAY == TCG_LAY_KM) { rc = encode_f32(&a.mapB, a.b.x, a.R, a.Q, a.ldb, BN); a.b_tx = BN * 128; }
    else                    { rc = encode_f32(&a.mapB, a.b.x, a.Q, a.R, a.ldb, 32); a.b_tx = a.b_atoms * 4096; }
    if (rc) return rc;
    const int total = numPt * numQt * a.splits;
    const int grid = total < NUM_SMS ? total : NUM_SMS;
    { cudaError_t le = b200sp_launch_pdl(a.groups > 1 ? kern_g : kern, dim3(grid), dim3(NT), smem, st, a); if (le != cudaSuccess) return (int)le; }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

// 16 epilogue warps need 37 KB more shared memory (staging tiles + statistics rows) than 8: taken only when that costs neither
// tile width nor the resident weights (r3e: shapes that lost either ran 20-50 % slower than with 8 warps)
template <int ALAY, int BLAY, int EPI, int AMODE, int BMODE>
int launch_ew(Tcg2Args& a, cudaStream_t st, bool want16, bool force16, bool lean = false) {
    // measured on the whole step (r3n / r3o): 8 + 8 warps 5.66 ms; 8 + 4 warps (144 registers) 5.71; 8 + 16 (72 registers) 5.78;
    // the weight-gradient kernel (main-loop-bound) loses 0.06 ms with 8 converter warps and keeps 16
    if (lean) return launch_cfg<ALAY, BLAY, EPI, AMODE, BMODE, 8, 8>(a, st);
    if (want16 && !force16) {
        Tcg2Args a8 = a, a16 = a;
        const int r8 = launch_cfg<ALAY, BLAY, EPI, AMODE, BMODE, 8>(a8, st, true);
        const int r16 = launch_cfg<ALAY, BLAY, EPI, AMODE, BMODE, 16>(a16, st, true);
        want16 = r8 == 0 && r16 == 0 && a16.BN == a8.BN && a16.b_res == a8.b_res;
    }
    if (want16) return launch_cfg<ALAY, BLAY, EPI, AMODE, BMODE, 16>(a, st);
    return launch_cfg<ALAY, BLAY, EPI, AMODE, BMODE, 8>(a, st);
}

constexpr int PRE_NT = 8 * 32 + MMA_T + TMA_T;

// planes [nm][rows][ld] of a presplit operand: {R, rows, nm} with a {32, box_rows, nm} box, 128-byte swizzle, zero fill
inline int encode_planes(CUtensorMap* m, const float* base, int R, int rows, int ld, size_t plane, int nm, int box_rows) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return B200SP_ENOSYS;
    cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)rows, (cuuint64_t)nm};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(nm > 1 ? plane : (size_t)rows * ld) * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)nm};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : B200SP_ENOSYS;
}

// cycles of one k-block of a 128 x BN tile in PRE mode (MMA issue + tensor pipe, r2 timelines) and of a tile's fixed part
inline long long pre_cost(int tiles, int kb_per_tile, int BN) {
    const long long waves = ceil_div(tiles, NUM_SMS);
    return waves * ((long long)kb_per_tile * (8 * BN + 250) + 40 * BN + 1500);
}

template <int EPI>
int launch_pre(Tcg2Args& a, const float* A, int lda, size_t planeA, const float* B, int ldb, size_t planeB, cudaStream_t st) {
    using E = ETf;
    constexpr int EPI_W = 8, EPI_SETS = 2;
    auto kern = tcgemm2_kernel<TCG_LAY_KM, TCG_LAY_KM, EPI, XM_PLAIN, XM_PLAIN, false, true, 8, 8>;       // launch bound 576 threads (320 launched): 112 registers
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int numPt = ceil_div(a.P, BM);
    a.nkb = ceil_div(a.R, E::KE);
    {
        const int rem = a.R - (a.nkb - 1) * E::KE;
        a.nks_last = ceil_div(rem, 2 * E::EPV);
        a.ac_last = 2 * a.nks_last;
    }
    // ---- tile width and (weight gradient) split-K factor: fewest waves x cycles per wave ----
    int bestBN = 0, bestS = 1;
    long long best = -1;
    static int cap_env = -1;
    if (cap_env < 0) { const char* e = getenv("B200SP_TCG2_PRE_BN"); cap_env = e ? atoi(e) : 0; }      // experiments: force the tile width cap
    for (int cap = cap_env ? cap_env : 128; cap >= (cap_env ? cap_env : 32); cap -= 16) {
        const int nq = ceil_div(a.Q, cap);
        const int BN = ceil_div(ceil_div(a.Q, nq), 16) * 16;
        if (BN > cap) continue;
        const int smax = EPI == TCG_EPI_ATOMIC ? (a.nkb / 4 > 0 ? (a.nkb / 4 < 8 ? a.nkb / 4 : 8) : 1) : 1;
        for (int sp = 1; sp <= smax; ++sp) {
            const int kps = ceil_div(a.nkb, sp);
            if (ceil_div(a.nkb, kps) != sp) continue;
            const long long c = pre_cost(numPt * nq * sp, kps, BN) + (sp > 1 ? 300 : 0);
            if (best < 0 || c < best) { best = c; bestBN = BN; bestS = sp; }
        }
    }
    const int BN = bestBN, numQt = ceil_div(a.Q, BN);
    a.BN = BN; a.numPt = numPt; a.numQt = numQt;
    a.splits = bestS;
    a.kb_per_split = ceil_div(a.nkb, a.splits);
    a.a_op_bytes = BM * 128;
    a.b_op_bytes = BN * 128;
    a.a_tx = a.a_op_bytes; a.b_tx = a.b_op_bytes;
    a.b_res = 0; a.ts = 0; a.groups = 1; a.n_raw = 1; a.raw_stage_bytes = 0;
    a.op_stage_bytes = a.nm * (a.a_op_bytes + a.b_op_bytes);
    const uint32_t fixed = EPI_W * 32 * STG_LD * 4 + EPI_W * 2 * BN * 4 + 512 + 1088 + 1024;
    a.n_op = MAX_OP;
    while (a.n_op > 1 && a.n_op * a.op_stage_bytes + fixed > SMEM_LIMIT) --a.n_op;
    if (a.n_op < 2) return B200SP_ENOSYS;
    a.off_bres = a.n_op * a.op_stage_bytes;
    a.off_raw = a.off_bres;
    a.off_stg = a.off_raw;
    a.off_stat = a.off_stg + EPI_W * 32 * STG_LD * 4;
    a.off_bar = (a.off_stat + EPI_W * 2 * BN * 4 + 15) & ~15u;
    const uint32_t smem = a.off_bar + 512 + 1024;
    a.acc_cols = (a.nm == 2 && 2 * 2 * BN <= 512) ? 2 * BN : BN;
    a.nacc = 4 * a.acc_cols <= 512 ? 4 : 2;
    a.stack_b = (a.nm == 2 && a.acc_cols == 2 * BN && 2 * BN <= 256) ? 1 : 0;
    a.split_epi = (BN <= 32 && numQt == 1 && EPI_SETS >= 2) ? 1 : 0;
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.nacc * a.acc_cols)) cols <<= 1;
    a.tmem_cols = cols;
    int rc = encode_planes(&a.mapA, A, a.R, a.P, lda, planeA, a.nm, BM);
    if (rc) return rc;
    rc = encode_planes(&a.mapB, B, a.R, a.Q, ldb, planeB, a.nm, BN);
    if (rc) return rc;
    a.mapA2 = a.mapA;
    const int total = numPt * numQt * a.splits;
    const int grid = total < NUM_SMS ? total : NUM_SMS;
    { cudaError_t le = b200sp_launch_pdl(kern, dim3(grid), dim3(PRE_NT), smem, st, a); if (le != cudaSuccess) return (int)le; }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

float* g_ws_base = nullptr;
size_t g_ws_bytes = 0;

}  // namespace

extern "C" int b200sp_set_workspace(void* base, size_t bytes) {
    if (((uintptr_t)base & 255) != 0) return B200SP_EINVAL;
    g_ws_base = reinterpret_cast<float*>(base);
    g_ws_bytes = base ? bytes : 0;
    return 0;
}
void tcg_workspace(float** base, size_t* bytes) { *base = g_ws_base; *bytes = g_ws_bytes; }

// Presplit route for the tensor-bound, L2-resident layers: reduction-side rows <= B200SP_TCG2_PRE_M (default 4096: the 7x7 layers
// at the benchmark batch), both other extents at least 64.  See opsplit.cu for why.
int tcgemm2_presplit(const TcgProblem& p, cudaStream_t st) {
    static int on = -1, max_m = 4096;
    if (on < 0) {
        const char* e = getenv("B200SP_TCG2_PRE");
        on = (e && e[0] == '0') ? 0 : 1;
        const char* m = getenv("B200SP_TCG2_PRE_M");
        if (m) max_m = atoi(m);
    }
    if (!on || !g_ws_base) return B200SP_ENOSYS;
    if (p.dtype != B200SP_F32 && p.dtype != B200SP_F32_TF32X1) return B200SP_ENOSYS;
    if (p.b.mode == B200SP_VT_DY) return B200SP_ENOSYS;
    if ((p.a.mode == B200SP_VT_BNACT && p.a.act == B200SP_ACT_SIGMOID) || (p.b.mode == B200SP_VT_BNACT && p.b.act == B200SP_ACT_SIGMOID)) return B200SP_ENOSYS;
    // M of the layer: rows of the activation operand (P in fwd / dgrad, the reduction in wgrad)
    const int m_layer = p.epi == TCG_EPI_ATOMIC ? p.R : p.P;
    if (m_layer > max_m || m_layer < 256 || p.R < 128 || p.Q < 64 || p.P < 64) return B200SP_ENOSYS;
    if (p.P % 4 || p.Q % 4 || p.R % 4 || p.lda % 4 || p.ldb % 4) return B200SP_ENOSYS;
    const int ldo = p.ldo > 0 ? p.ldo : p.Q;
    if (ldo % 4) return B200SP_ENOSYS;
    if (((uintptr_t)p.a.x | (uintptr_t)p.b.x | (uintptr_t)p.a.x2 | (uintptr_t)p.out) & 15) return B200SP_ENOSYS;
    const int nm = p.dtype == B200SP_F32_TF32X1 ? 1 : 2;
    // planes [nm][P][R] and [nm][Q][R] (R padded to a multiple of 4 floats = the 16-byte stride granularity of a tensor map)
    const int ldk = (p.R + 3) & ~3;
    const size_t planeA = ((size_t)p.P * ldk + 63) & ~(size_t)63, planeB = ((size_t)p.Q * ldk + 63) & ~(size_t)63;
    // the engines run the weight gradients on a second stream, concurrently with the data gradients: each kind has its own half
    const size_t half = (g_ws_bytes / 2) & ~(size_t)255;
    if ((nm * (planeA + planeB)) * sizeof(float) > half) return B200SP_ENOSYS;
    float* wa = g_ws_base + (p.epi == TCG_EPI_ATOMIC ? half / sizeof(float) : 0);
    float* wb = wa + nm * planeA;
    OpSplitJob ja, jb;
    memset(&ja, 0, sizeof(ja)); memset(&jb, 0, sizeof(jb));
    ja.t = p.a; jb.t = p.b;
    ja.trans = p.a_lay == TCG_LAY_MM; jb.trans = p.b_lay == TCG_LAY_MM;
    // stored shapes: K-major [MN][R]; MN-major [R][MN]
    ja.rows = ja.trans ? p.R : p.P; ja.cols = ja.trans ? p.P : p.R; ja.ld = p.lda;
    jb.rows = jb.trans ? p.R : p.Q; jb.cols = jb.trans ? p.Q : p.R; jb.ld = p.ldb;
    ja.out = wa; ja.out_ld = ldk; ja.plane = planeA; ja.nm = nm;
    jb.out = wb; jb.out_ld = ldk; jb.plane = planeB; jb.nm = nm;
    int rc = opsplit_launch(ja, jb, st);
    if (rc) return rc;
    Tcg2Args a;
    memset(&a, 0, sizeof(a));
    a.a = p.a; a.b = p.b;
    a.a.mode = a.b.mode = B200SP_VT_PLAIN;
    a.P = p.P; a.Q = p.Q; a.R = p.R; a.lda = ldk; a.ldb = ldk;
    a.ldo = ldo;
    a.out = p.out; a.bias = p.bias; a.out_act = p.out_act;
    a.has_bnf = p.bnf != nullptr;
    if (p.bnf) a.bnf = *p.bnf;
    a.skip = p.skip; a.scale_out = p.scale_out;
    a.has_bnb = p.bnb != nullptr;
    if (p.bnb) a.bnb = *p.bnb;
    a.count = p.count;
    a.epi_sleep = 512;
    a.nm = nm;
    if (p.epi == TCG_EPI_FWD) return launch_pre<TCG_EPI_FWD>(a, wa, ldk, planeA, wb, ldk, planeB, st);
    if (p.epi == TCG_EPI_DGRAD) return launch_pre<TCG_EPI_DGRAD>(a, wa, ldk, planeA, wb, ldk, planeB, st);
    return launch_pre<TCG_EPI_ATOMIC>(a, wa, ldk, planeA, wb, ldk, planeB, st);
}

#ifdef TCG_TIMELINE
extern "C" int b200sp_tcg2_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_tcg2_tl, sizeof(long long) * 14 * 512);
}
#endif

int tcgemm2_launch(const TcgProblem& p, cudaStream_t st) {
    if (p.dtype != B200SP_F32 && p.dtype != B200SP_F32_TF32X1) return B200SP_ENOSYS;
    {   // tensor-bound L2-resident layers: operands split once by a pre-pass, TMA -> MMA with no converters
        const int rc = tcgemm2_presplit(p, st);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (p.Q % 4 || p.lda % 4 || p.ldb % 4) return B200SP_ENOSYS;
    if (p.a_lay == TCG_LAY_KM ? (p.R % 4) : (p.P % 4)) return B200SP_ENOSYS;
    if (p.b_lay == TCG_LAY_KM ? (p.R % 4) : (p.Q % 4)) return B200SP_ENOSYS;
    if (p.b.mode == B200SP_VT_DY) return B200SP_ENOSYS;
    if (((uintptr_t)p.a.x | (uintptr_t)p.b.x | (uintptr_t)p.a.x2 | (uintptr_t)p.out) & 15) return B200SP_ENOSYS;
    if ((p.a.mode == B200SP_VT_BNACT && p.a.act == B200SP_ACT_SIGMOID) || (p.b.mode == B200SP_VT_BNACT && p.b.act == B200SP_ACT_SIGMOID)) return B200SP_ENOSYS;
    Tcg2Args a;
    memset(&a, 0, sizeof(a));
    a.a = p.a; a.b = p.b;
    a.P = p.P; a.Q = p.Q; a.R = p.R; a.lda = p.lda; a.ldb = p.ldb;
    a.ldo = p.ldo > 0 ? p.ldo : p.Q;
    if (a.ldo % 4) return B200SP_ENOSYS;
    a.out = p.out; a.bias = p.bias; a.out_act = p.out_act;
    a.has_bnf = p.bnf != nullptr;
    if (p.bnf) a.bnf = *p.bnf;
    a.skip = p.skip; a.scale_out = p.scale_out;
    a.has_bnb = p.bnb != nullptr;
    if (p.bnb) a.bnb = *p.bnb;
    a.count = p.count;
    a.epi_sleep = 512;
    a.nm = p.dtype == B200SP_F32_TF32X1 ? 1 : 2;
    const int am = p.a.mode, bm = p.b.mode;
    // short reduction (<= B200SP_TCG2_EW_R, default 32 = one k-block): the epilogue bounds the kernel -> 16 epilogue warps
    // (r3f: +15 % on fwd [602112,96,16], +23 % on dgrad [602112,16,32]; longer reductions lose 5-15 % to the halved converter)
    static int ew_env = -1, ew_r = 32;
    if (ew_env < 0) {
        const char* e = getenv("B200SP_TCG2_EW");
        ew_env = e ? atoi(e) : 0;                   // 0 auto | 8 | 16
        const char* r = getenv("B200SP_TCG2_EW_R");
        if (r) ew_r = atoi(r);
    }
    const bool ew16 = ew_env == 16 || (ew_env == 0 && p.R <= ew_r);
    // 576-thread variant (8 epilogue + 8 converter warps, 112 registers per thread): B200SP_TCG2_LEAN = dgrad | all, reduction <= LEAN_R
    static int lean_mode = -1, lean_r = 1 << 30;
    if (lean_mode < 0) {
        const char* e = getenv("B200SP_TCG2_LEAN");
        lean_mode = !e ? 2 : (e[0] == 'd' ? 1 : (e[0] == 'a' ? 2 : 0));       // default: all (r3n: KRN step 5.78 -> 5.66 ms)
        const char* r = getenv("B200SP_TCG2_LEAN_R");
        if (r) lean_r = atoi(r);
    }
    const bool lean_d = lean_mode >= 1 && p.R <= lean_r, lean_f = lean_mode == 2 && p.R <= lean_r;
    if (p.epi == TCG_EPI_FWD && p.a_lay == TCG_LAY_KM && p.b_lay == TCG_LAY_KM && bm == B200SP_VT_PLAIN) {
        if (am == B200SP_VT_PLAIN) return launch_ew<TCG_LAY_KM, TCG_LAY_KM, TCG_EPI_FWD, XM_PLAIN, XM_PLAIN>(a, st, ew16, ew_env == 16, lean_f);
        if (am == B200SP_VT_BNACT) return launch_ew<TCG_LAY_KM, TCG_LAY_KM, TCG_EPI_FWD, XM_BNACT, XM_PLAIN>(a, st, ew16, ew_env == 16, lean_f);
    }
    if (p.epi == TCG_EPI_DGRAD && p.a_lay == TCG_LAY_KM && p.b_lay == TCG_LAY_MM && bm == B200SP_VT_PLAIN) {
        if (am == B200SP_VT_PLAIN) return launch_ew<TCG_LAY_KM, TCG_LAY_MM, TCG_EPI_DGRAD, XM_PLAIN, XM_PLAIN>(a, st, ew16, ew_env == 16, lean_d);
        if (am == B200SP_VT_DY) return launch_ew<TCG_LAY_KM, TCG_LAY_MM, TCG_EPI_DGRAD, XM_DY, XM_PLAIN>(a, st, ew16, ew_env == 16, lean_d);
    }
    if (p.epi == TCG_EPI_ATOMIC && p.a_lay == TCG_LAY_MM && p.b_lay == TCG_LAY_MM) {
        if (am == B200SP_VT_PLAIN && bm == B200SP_VT_PLAIN) return launch_cfg<TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_PLAIN, XM_PLAIN>(a, st);
        if (am == B200SP_VT_PLAIN && bm == B200SP_VT_BNACT) return launch_cfg<TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_PLAIN, XM_BNACT>(a, st);
        if (am == B200SP_VT_DY && bm == B200SP_VT_PLAIN) return launch_cfg<TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_DY, XM_PLAIN>(a, st);
        if (am == B200SP_VT_DY && bm == B200SP_VT_BNACT) return launch_cfg<TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_DY, XM_BNACT>(a, st);
    }
    return B200SP_ENOSYS;
}
