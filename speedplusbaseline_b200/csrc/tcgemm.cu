// tcgemm.cu -- persistent, warp-specialised tcgen05 GEMM for the 1x1-convolution / FC family
// (forward, dgrad, wgrad) with the BatchNorm/activation transforms fused on BOTH sides:
//
//   producers (8 warps) : cp.async raw 16-byte pieces of the A and B operands into thread-private
//                         staging slots (deep prefetch, no register staging), then apply the
//                         virtual-tensor transform (BN affine + activation, or the folded BatchNorm
//                         backward dy = cA*g + cB*y + cC), split fp32 into (tf32 hi, lo) or round to
//                         bf16, and store into the 128B-swizzled UMMA operand tiles;
//   MMA warp (1 thread) : tcgen05.mma kind::tf32 (x3: lo*hi + hi*lo + hi*hi, fp32-grade accuracy) or
//                         kind::f16 (bf16) with the 128 x BN fp32 accumulator in TMEM, double-buffered
//                         so the epilogue of tile i overlaps the main loop of tile i+1;
//   epilogue (4 warps)  : tcgen05.ld -> per-warp smem transpose -> coalesced global phase doing
//                         bias/activation (fwd), skip-add + activation-derivative mask (dgrad) or
//                         fp32 red.add (wgrad), plus the per-channel BatchNorm reductions
//                         (sum, sumsq | sum g, sum g*xhat) accumulated per CTA and flushed with one
//                         double atomic per channel; the last CTA finalises the statistics.
//
// Operands are consumed in their natural row-major layout: K-major when the reduction dimension is
// contiguous (fwd A/B, dgrad A), MN-major otherwise (dgrad B = W, wgrad A = dY and B = X), so no
// transposed weight copies exist anywhere.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "tcgemm.cuh"
#include "tc_common.cuh"

#ifdef TCG_TIMELINE
__device__ long long g_tcg_tl[8][256];
#define TL(row, idx) do { if (blockIdx.x == 0 && (idx) < 256) g_tcg_tl[row][idx] = clock64(); } while (0)
#else
#define TL(row, idx) do {} while (0)
#endif

namespace {

constexpr int EPI_T = 256, MMA_T = 32, PROD_T = 512;
constexpr int APT = 1024 / PROD_T;              // 16-byte pieces of a 128-row x 128-byte tile per producer thread
constexpr int EPI_W = EPI_T / 32;               // epilogue warps: warp w reads TMEM lane quarter w%4; with 8 warps the column chunks are split by parity w/4
constexpr int EPI_SETS = EPI_W / 4;
constexpr int MMA_WARP = EPI_W;
constexpr int NT = EPI_T + MMA_T + PROD_T;     // 672 threads: warps 0-3 epilogue, 4 MMA, 5-20 producers (the operand transform is latency-bound: it needs the warps)
constexpr int BM = 128;
constexpr int STG_LD = 36;                      // floats per epilogue staging row (32 + pad, 16B aligned)
constexpr int MAX_OP = 4, MAX_RAW = 4;

template <typename T> struct ET;
template <> struct ET<float> { static constexpr int ES = 4, KE = 32, EPV = 4, NM = 2; static constexpr bool TF32 = true; };
template <> struct ET<bf16>  { static constexpr int ES = 2, KE = 64, EPV = 8, NM = 1; static constexpr bool TF32 = false; };

struct TcgArgs {
    b200sp_vtensor a, b;
    int P, Q, R, lda, ldb, ldo;
    int BN, numPt, numQt, splits, kb_per_split, nkb;
    int n_op, n_raw;
    int nb_pieces, nb_slots;         // B pieces per k-block; per-thread B slots = ceil(nb_pieces / PROD_T)
    int a_dy;
    int b_res;                       // B (weights) converted once per CTA and kept resident (numQt == 1, fits)
    int ac_last, nks_last;           // K-major operands: active 16-byte chunks / UMMA k-steps of the LAST k-block
    uint32_t off_bres;
    uint32_t a_op_bytes, b_op_bytes; // bytes of ONE math copy (hi or lo) of each operand tile
    uint32_t op_stage_bytes, raw_stage_bytes;
    uint32_t off_raw, off_stg, off_stat, off_bar;   // dynamic smem carve-up (from the 1024-aligned base)
    uint32_t tmem_cols;
    int nacc;                        // TMEM accumulator ring depth (2 or 4)
    int acc_cols;                    // TMEM columns per ring slot: BN, or 2*BN when the 3xTF32 correction terms have their own accumulator
    int split_epi;                   // one column chunk per tile: the two epilogue warp sets take alternate tiles
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
    uint32_t epi_sleep;              // epilogue back-off in ns (B200SP_LEAN_TCG builds only)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_T) : "memory"); }

// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity, int mode = 0) {
    uint32_t spins = 0;
    if (mode == 0) {
        while (!tc::mbar_try_wait_hint(bar, parity, 20000u)) {      // hardware-suspended wait (no issue slots burnt)
            if (++spins > (1u << 22)) __trap();
        }
    } else if (mode == 1) {
        while (!tc::mbar_try_wait(bar, parity)) {
            if (++spins > (1u << 26)) __trap();
        }
    } else {
        while (!tc::mbar_test_wait(bar, parity)) {
            if (++spins > (1u << 28)) __trap();
        }
    }
}

// long waits (the epilogue waits for a whole main loop): back off so the spin does not steal issue
// slots from the producer warps sharing the scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
    uint32_t spins = 0;
    while (!tc::mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++spins > (1u << 24)) __trap();
    }
}

// ---- work decomposition ---------------------------------------------------------------------------
struct Item {
    int p0, q0, kb0, kb1;
};
__device__ __forceinline__ Item get_item(const TcgArgs& g, int it) {
    Item w;
    const int qt = it % g.numQt;
    const int t2 = it / g.numQt;
    const int pt = t2 % g.numPt, sp = t2 / g.numPt;
    w.p0 = pt * BM;
    w.q0 = qt * g.BN;
    w.kb0 = sp * g.kb_per_split;
    w.kb1 = min(g.nkb, w.kb0 + g.kb_per_split);
    return w;
}
// flattened (item, k-block) iterator over this CTA's work
struct KIter {
    int it, total, stride;
    Item w;
    int kb;
    __device__ __forceinline__ void init(const TcgArgs& g, int first, int total_, int stride_) {
        it = first; total = total_; stride = stride_;
        if (it < total) { w = get_item(g, it); kb = w.kb0; }
    }
    __device__ __forceinline__ bool valid() const { return it < total; }
    __device__ __forceinline__ void next(const TcgArgs& g) {
        if (++kb >= w.kb1) {
            it += stride;
            if (it < total) { w = get_item(g, it); kb = w.kb0; }
        }
    }
};

// ---- shared-memory / conversion helpers -----------------------------------------------------------
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---- per-channel transform of a virtual tensor, mode fixed at compile time --------------------------
enum { XM_PLAIN = 0, XM_BNACT = 1, XM_DY = 2 };
struct XfP { float4 a, b, c; };             // BNACT: scale, shift      DY: cA, cB, cC   (4 channels)
template <int MODE>
__device__ __forceinline__ void xf_load(const b200sp_vtensor& t, int ch, XfP& p) {
    if (MODE != XM_PLAIN) { p.a = ldg4(t.p0 + ch); p.b = ldg4(t.p1 + ch); }
    if (MODE == XM_DY) p.c = ldg4(t.p2 + ch);
}
template <int MODE>
__device__ __forceinline__ float4 xf_apply(float4 x, float4 x2, const XfP& p, ActP act) {
    if (MODE == XM_PLAIN) return x;
    if (MODE == XM_BNACT) {
#ifdef B200SP_LEAN_TCG
        // ReLU / ReLU6 (slope 0; hi = +inf or 6): min(max(z, 0), hi) is two instructions per value instead of the four of the
        // branch-free generic form, and identical for finite z.  The branch is kernel-uniform.
        if (act.slope == 0.f)
            return make_float4(fminf(fmaxf(fmaf(x.x, p.a.x, p.b.x), 0.f), act.hi), fminf(fmaxf(fmaf(x.y, p.a.y, p.b.y), 0.f), act.hi),
                               fminf(fmaxf(fmaf(x.z, p.a.z, p.b.z), 0.f), act.hi), fminf(fmaxf(fmaf(x.w, p.a.w, p.b.w), 0.f), act.hi));
#endif
        return make_float4(act_fwd(fmaf(x.x, p.a.x, p.b.x), act), act_fwd(fmaf(x.y, p.a.y, p.b.y), act),
                           act_fwd(fmaf(x.z, p.a.z, p.b.z), act), act_fwd(fmaf(x.w, p.a.w, p.b.w), act));
    }
    return make_float4(fmaf(p.a.x, x.x, fmaf(p.b.x, x2.x, p.c.x)), fmaf(p.a.y, x.y, fmaf(p.b.y, x2.y, p.c.y)),
                       fmaf(p.a.z, x.z, fmaf(p.b.z, x2.z, p.c.z)), fmaf(p.a.w, x.w, fmaf(p.b.w, x2.w, p.c.w)));
}

// ---- operand loader: one producer thread's share of an operand tile ---------------------------------
// The 16-byte pieces of a tile are dealt to the 512 producer threads so that each thread owns a fixed
// (row, chunk) pattern; piece i (i < np) lives at operand-smem offset soff + i*8192 and in this thread's
// private raw slot i.  Everything that depends only on the TILE (global offsets, validity bits) is computed
// once per tile by begin(); per k-block the loader adds kb*kstride -- the main loop carries no index math.
//   K-major : chunk gc = pt & 7 of row  (pt >> 3) + 64 i;  channel = reduction index = kb*KE + gc*EPV,
//             the same for all of a thread's pieces: parameters are fetched once per k-block (early, so
//             the load overlaps the barrier waits).
//   MN-major: chunk gc of reduction row (pt >> 3) & (KE-1);  M/N atom 2i + (pt >> 8) for tf32 (32-row
//             k-blocks), atom i for bf16;  channel = M/N index, fixed per piece for a whole tile.
template <typename T, int LAY, int MODE, int NP>
struct OpLoader {
    using E = ET<T>;
    static constexpr int MNA = 128 / E::ES;                                   // M/N elements per 128-byte atom
    static constexpr int ASTEP = (LAY == TCG_LAY_MM && E::TF32) ? 2 : 1;      // atoms between a thread's pieces
    struct FTile { size_t off[NP]; uint32_t ok; };                            // fetch side
    struct CTile { uint32_t ok; int ch0; };                                   // convert side
    b200sp_vtensor vt;
    const char *x, *x2;
    ActP act;
    int ld, mn_ext, R, nkb, ac_last;
    int gc, row, abase, np;
    uint32_t soff;
    size_t kstride;
    XfP par, par2;

    __device__ __forceinline__ void init(const b200sp_vtensor& t, int ld_, int mn_ext_, int R_, int nkb_, int ac_last_, int pt,
                                         int tile_rows) {
        vt = t; x = reinterpret_cast<const char*>(t.x); x2 = reinterpret_cast<const char*>(t.x2);
        act = act_params(t.act);
        ld = ld_; mn_ext = mn_ext_; R = R_; nkb = nkb_; ac_last = ac_last_;
        gc = pt & 7;
        if (LAY == TCG_LAY_KM) {
            row = pt >> 3; abase = 0;
            soff = tc::sw128_off(row, gc);
            kstride = 128;
            const int rem = tile_rows - row;
            np = rem <= 0 ? 0 : (rem + 63) >> 6;
        } else {
            const int atoms = (tile_rows * E::ES + 127) / 128;
            if (E::TF32) {
                row = (pt >> 3) & 31; abase = pt >> 8;
                soff = tc::sw128b32_off(row, gc) + abase * 4096;
                np = (atoms - abase + 1) >> 1;
            } else {
                row = pt >> 3; abase = 0;
                soff = tc::sw128_off(row, gc);
                np = atoms;
            }
            kstride = (size_t)E::KE * ld * E::ES;
        }
        if (np > NP) np = NP;
        if (np < 0) np = 0;
    }
    __device__ __forceinline__ void begin(FTile& f, int mn0) const {
        f.ok = 0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (LAY == TCG_LAY_KM) {
                const int mn = mn0 + row + 64 * i;
                f.off[i] = ((size_t)mn * ld + gc * E::EPV) * E::ES;
                if (mn < mn_ext) f.ok |= 1u << i;
            } else {
                const int mn = mn0 + (ASTEP * i + abase) * MNA + gc * E::EPV;
                f.off[i] = ((size_t)row * ld + mn) * E::ES;
                if (mn < mn_ext) f.ok |= 1u << i;
            }
        }
    }
    __device__ __forceinline__ void begin(CTile& c, int mn0) const {
        c.ok = 0;
        c.ch0 = mn0 + abase * MNA + gc * E::EPV;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int mn = LAY == TCG_LAY_KM ? mn0 + row + 64 * i : c.ch0 + ASTEP * MNA * i;
            if (mn < mn_ext) c.ok |= 1u << i;
        }
    }
    // does this thread have anything to do for k-block kb?  (K-major: the partial last k-block only holds
    // ac_last chunks;  MN-major: reduction rows beyond R are zero-filled)
    __device__ __forceinline__ bool active(int kb) const {
        return LAY == TCG_LAY_KM ? (kb != nkb - 1 || gc < ac_last) : true;
    }
    __device__ __forceinline__ bool kb_ok(int kb) const {
        return LAY == TCG_LAY_KM ? (kb * E::KE + gc * E::EPV < R) : (kb * E::KE + row < R);
    }
    // cp.async this thread's raw pieces of k-block kb into its private slots (slot i at raw + i*PROD_T*16)
    __device__ __forceinline__ void issue(int kb, const FTile& f, uint32_t raw, uint32_t raw2) const {
        if (!active(kb)) return;
        const bool kok = kb_ok(kb);
        const size_t kofs = (size_t)kb * kstride;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < np) {
                const bool ok = kok && ((f.ok >> i) & 1u);
                const size_t o = ok ? f.off[i] + kofs : 0;
                cp_async16(raw + i * (PROD_T * 16), x + o, ok);
                if (MODE == XM_DY) cp_async16(raw2 + i * (PROD_T * 16), x2 + o, ok);
            }
        }
    }
    // K-major: fetch the transform parameters of k-block kb (call early: the latency hides behind the waits)
    __device__ __forceinline__ void load_params(int kb) {
        if (LAY == TCG_LAY_KM && MODE != XM_PLAIN) {
            int ch = kb * E::KE + gc * E::EPV;
            ch = min(ch, R - (E::TF32 ? 4 : 8));          // columns beyond R meet zero-filled B columns: any finite value does
            xf_load<MODE>(vt, ch, par);
            if (!E::TF32) xf_load<MODE>(vt, ch + 4, par2);
        }
    }
    // transform the raw pieces and store them into the operand tile (hi [, lo])
    __device__ __forceinline__ void convert(int kb, const CTile& c, uint32_t raw, uint32_t raw2, uint32_t op_hi, uint32_t op_lo) {
        if (!active(kb)) return;
        const bool kok = kb_ok(kb);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < np) {
                const uint32_t so = soff + i * 8192;
                float4 r = lds4(raw + i * (PROD_T * 16)), r2 = f4zero();
                if (MODE == XM_DY) r2 = lds4(raw2 + i * (PROD_T * 16));
                // zero-filled raw data stays zero only for PLAIN; transformed operands must be masked where the
                // OTHER operand is not guaranteed to be zero: MN-major reduction rows / columns beyond the tensor
                bool ok = true;
                if (LAY == TCG_LAY_MM && MODE != XM_PLAIN) {
                    ok = kok && ((c.ok >> i) & 1u);
                    int ch = c.ch0 + ASTEP * MNA * i;
                    ch = min(ch, mn_ext - (E::TF32 ? 4 : 8));
                    xf_load<MODE>(vt, ch, par);
                    if (!E::TF32) xf_load<MODE>(vt, ch + 4, par2);
                }
                if (E::TF32) {
                    float4 v = xf_apply<MODE>(r, r2, par, act);
                    if (!ok) v = f4zero();
                    float4 h, l;
                    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
                    tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
                    sts4(op_hi + so, h);
                    sts4(op_lo + so, l);
                } else {
                    float4 o = r;
                    if (MODE != XM_PLAIN) {
                        const uint32_t* u = reinterpret_cast<const uint32_t*>(&r);
                        const uint32_t* u2 = reinterpret_cast<const uint32_t*>(&r2);
                        const float2 a0 = bf2_to_f2(u[0]), a1 = bf2_to_f2(u[1]), a2 = bf2_to_f2(u[2]), a3 = bf2_to_f2(u[3]);
                        const float2 b0 = bf2_to_f2(u2[0]), b1 = bf2_to_f2(u2[1]), b2 = bf2_to_f2(u2[2]), b3 = bf2_to_f2(u2[3]);
                        const float4 lo4 = xf_apply<MODE>(make_float4(a0.x, a0.y, a1.x, a1.y), make_float4(b0.x, b0.y, b1.x, b1.y), par, act);
                        const float4 hi4 = xf_apply<MODE>(make_float4(a2.x, a2.y, a3.x, a3.y), make_float4(b2.x, b2.y, b3.x, b3.y), par2, act);
                        uint32_t* w = reinterpret_cast<uint32_t*>(&o);
                        w[0] = f2_to_bf2(lo4.x, lo4.y); w[1] = f2_to_bf2(lo4.z, lo4.w);
                        w[2] = f2_to_bf2(hi4.x, hi4.y); w[3] = f2_to_bf2(hi4.z, hi4.w);
                    }
                    if (!ok) o = f4zero();
                    sts4(op_hi + so, o);
                }
            }
        }
    }
};

// =====================================================================================================
template <typename T, int ALAY, int BLAY, int EPI, int AMODE, int BMODE>
__global__ void __launch_bounds__(NT, 1) tcgemm_kernel(const TcgArgs g) {
    using E = ET<T>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = tc::smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
    uint64_t* full = bars;                    // [MAX_OP]  producers -> MMA
    uint64_t* empty = bars + MAX_OP;          // [MAX_OP]  MMA -> producers
    uint64_t* tfull = bars + 2 * MAX_OP;      // [4]       MMA -> epilogue
    uint64_t* tempty = bars + 2 * MAX_OP + 4; // [4]       epilogue -> MMA
    uint64_t* bfull = bars + 2 * MAX_OP + 8;  // [1]       producers -> MMA: resident B converted
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_OP + 9);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = g.numPt * g.numQt * g.splits;

    if (tid == 0) {
        for (int i = 0; i < MAX_OP; ++i) { tc::mbar_init(&full[i], PROD_T / 32); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], g.split_epi ? EPI_W / EPI_SETS : EPI_W); }
        tc::mbar_init(bfull, PROD_T / 32);
        tc::mbar_fence_init();
    }
    if (warp == MMA_WARP) { tc::tmem_alloc(tmem_slot, g.tmem_cols); tc::tmem_relinquish(); }
    if (tid < EPI_T) {          // zero the per-warp statistic accumulators
        float* st = reinterpret_cast<float*>(smem + g.off_stat);
        for (int i = tid; i < EPI_W * 2 * g.BN; i += EPI_T) st[i] = 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // broadcast form: TMEM addresses stay in uniform registers
    // operand tile offsets inside one ring stage (A only when B is resident)
    const uint32_t b_in_stage = E::NM * g.a_op_bytes;

    if (warp > MMA_WARP) {
        // ======================================= PRODUCERS =======================================
        const int pt = tid - (EPI_T + MMA_T);
        constexpr int NPB = (E::TF32 ? 1024 : 2048) / PROD_T;
        typedef OpLoader<T, ALAY, AMODE, APT> LoaderA;
        typedef OpLoader<T, BLAY, BMODE, NPB> LoaderB;
        LoaderA LA;
        LoaderB LB;
        LA.init(g.a, g.lda, g.P, g.R, g.nkb, g.ac_last, pt, BM);
        LB.init(g.b, g.ldb, g.Q, g.R, g.nkb, g.ac_last, pt, g.BN);
        const uint32_t raw0 = s_base + g.off_raw + pt * 16;
        constexpr int slotA2 = APT, slotB = AMODE == XM_DY ? 2 * APT : APT;
        typename LoaderA::FTile fa;
        typename LoaderA::CTile ca;
        typename LoaderB::FTile fb;
        typename LoaderB::CTile cb;
        if (g.b_res) {
            // weights: convert every k-block once, keep them resident for all of this CTA's tiles.  The raw
            // staging slots of the A ring double as a software pipeline so the loads of n_raw k-blocks overlap.
            LB.begin(fb, 0);
            LB.begin(cb, 0);
            for (int d = 0; d < g.n_raw; ++d) {
                if (d < g.nkb) LB.issue(d, fb, raw0 + d * g.raw_stage_bytes, 0);
                cp_async_commit();
            }
            int rsb = 0;
            for (int kb = 0; kb < g.nkb; ++kb) {
                if (g.n_raw == 4) cp_async_wait<3>(); else if (g.n_raw == 3) cp_async_wait<2>(); else cp_async_wait<1>();
                const uint32_t b_hi = s_base + g.off_bres + kb * (E::NM * g.b_op_bytes), b_lo = b_hi + g.b_op_bytes;
                LB.load_params(kb);
                LB.convert(kb, cb, raw0 + rsb * g.raw_stage_bytes, 0, b_hi, b_lo);
                if (kb + g.n_raw < g.nkb) LB.issue(kb + g.n_raw, fb, raw0 + rsb * g.raw_stage_bytes, 0);
                cp_async_commit();
                if (++rsb == g.n_raw) rsb = 0;
            }
            cp_async_wait<0>();
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bfull);
        }
        KIter fetch, cons;
        fetch.init(g, blockIdx.x, total, gridDim.x);
        cons.init(g, blockIdx.x, total, gridDim.x);
        int f_it = -1, c_it = -1;                  // work item whose tile state fa/fb (ca/cb) currently describes

        auto issue = [&](const KIter& k, int rs) {
            if (k.it != f_it) {
                LA.begin(fa, k.w.p0);
                if (!g.b_res) LB.begin(fb, k.w.q0);
                f_it = k.it;
            }
            const uint32_t rbase = raw0 + rs * g.raw_stage_bytes;
            LA.issue(k.kb, fa, rbase, rbase + slotA2 * (PROD_T * 16));
            if (!g.b_res) LB.issue(k.kb, fb, rbase + slotB * (PROD_T * 16), 0);
        };

        for (int d = 0; d < g.n_raw; ++d) {
            if (fetch.valid()) { issue(fetch, d); fetch.next(g); }
            cp_async_commit();
        }
        int rs = 0, os = 0, tln = 0;
        uint32_t par = 1;
        while (cons.valid()) {
            if (cons.it != c_it) {
                LA.begin(ca, cons.w.p0);
                if (!g.b_res) LB.begin(cb, cons.w.q0);
                c_it = cons.it;
            }
            LA.load_params(cons.kb);               // issued before the waits: their latency is hidden
            if (!g.b_res) LB.load_params(cons.kb);
            if (pt == 0) TL(0, tln);
            if (g.n_raw == 4) cp_async_wait<3>(); else if (g.n_raw == 3) cp_async_wait<2>(); else cp_async_wait<1>();
            if (pt == 0) TL(1, tln);
            mbar_wait_guard(&empty[os], par, g.wait_mode);
            if (pt == 0) TL(2, tln);
            const uint32_t rbase = raw0 + rs * g.raw_stage_bytes;
            const uint32_t a_hi = s_base + os * g.op_stage_bytes, a_lo = a_hi + g.a_op_bytes;
            const uint32_t b_hi = a_hi + b_in_stage, b_lo = b_hi + g.b_op_bytes;
            LA.convert(cons.kb, ca, rbase, rbase + slotA2 * (PROD_T * 16), a_hi, a_lo);
            if (!g.b_res) LB.convert(cons.kb, cb, rbase + slotB * (PROD_T * 16), 0, b_hi, b_lo);
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[os]);
            if (pt == 0) TL(3, tln);
            ++tln;
            // refill the raw slot just consumed with the k-block n_raw ahead
            if (fetch.valid()) { issue(fetch, rs); fetch.next(g); }
            cp_async_commit();
            cons.next(g);
            if (++rs == g.n_raw) rs = 0;
            if (++os == g.n_op) { os = 0; par ^= 1; }
        }
        cp_async_wait<0>();
    } else if (warp == MMA_WARP) {
        // ======================================= MMA ISSUER ======================================
        const uint32_t idesc = tc::make_idesc(E::TF32 ? tc::FMT_TF32 : tc::FMT_BF16, ALAY == TCG_LAY_MM, BLAY == TCG_LAY_MM, BM, g.BN);
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
        for (int it = blockIdx.x; it < total; it += gridDim.x) {
            const Item w = get_item(g, it);
            mbar_wait_guard(&tempty[acc], tpar, g.wait_mode);
            tc::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * g.acc_cols;
            // 3xTF32: the two correction products (lo*hi, hi*lo) go to a SECOND accumulator and are added in fp32-RN by the
            // epilogue.  tcgen05 accumulates with round-toward-zero, so every MMA into the large accumulator costs up to one
            // ulp of bias; keeping the small terms out of it cuts the number of such truncations from 3K/8 to K/8.
            const uint32_t d_corr = d_tmem + (g.acc_cols > g.BN ? g.BN : 0);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                if (lane == 0) TL(4, tlm);
                mbar_wait_guard(&full[os], fpar, g.wait_mode);
                tc::tc_fence_after();
                if (lane == 0) TL(5, tlm);
                if (tc::elect_one()) {
                    const uint32_t a_hi = s_base + os * g.op_stage_bytes, a_lo = a_hi + g.a_op_bytes;
                    const uint32_t b_hi = g.b_res ? s_base + g.off_bres + kb * (E::NM * g.b_op_bytes) : a_hi + b_in_stage;
                    const uint32_t b_lo = b_hi + g.b_op_bytes;
                    // a K-major operand only holds the active chunks of a partial last k-block: never read beyond them
                    const int nks = ((ALAY == TCG_LAY_KM || BLAY == TCG_LAY_KM) && kb == g.nkb - 1) ? g.nks_last : 4;
                    const uint32_t al = (a_lo >> 4) + A_LBO, ah = (a_hi >> 4) + A_LBO, bl = (b_lo >> 4) + B_LBO, bh = (b_hi >> 4) + B_LBO;
                    const uint32_t first = kb > w.kb0 ? 1u : 0u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (ks < nks) {
                            const uint32_t accum = ks > 0 ? 1u : first;
                            if (E::TF32) {
                                const bool split = g.acc_cols > g.BN;
                                tc::umma<true>(d_corr, mk(al + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                                tc::umma<true>(d_corr, mk(ah + ks * A_STEP, A_HIW), mk(bl + ks * B_STEP, B_HIW), idesc, 1u);
                                tc::umma<true>(d_tmem, mk(ah + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, split ? accum : 1u);
                            } else {
                                tc::umma<false>(d_tmem, mk(ah + ks * A_STEP, A_HIW), mk(bh + ks * B_STEP, B_HIW), idesc, accum);
                            }
                        }
                    }
                    tc::umma_commit(&empty[os]);
                    if (kb == w.kb1 - 1) tc::umma_commit(&tfull[acc]);
                    TL(6, tlm);
                }
                ++tlm;
                __syncwarp();
                if (++os == g.n_op) { os = 0; fpar ^= 1; }
            }
            if (++acc == g.nacc) { acc = 0; tpar ^= 1; }
        }
    } else {
        // ======================================= EPILOGUE ========================================
        const int lq = warp & 3, half = warp >> 2;      // TMEM lane quarter, column-chunk parity
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
        int last_chunk = -1;                            // last chunk this warp reads from TMEM
        const int c_first = g.split_epi ? 0 : half, c_step = g.split_epi ? 1 : EPI_SETS;
        for (int c = c_first; c < nchunks; c += c_step) last_chunk = c;
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
        uint32_t tpar = 1;
        for (int it = blockIdx.x; it < total; it += gridDim.x, ++ni) {
            if (++acc == g.nacc) acc = 0;
            if (acc == 0) tpar ^= 1;
            if (g.split_epi && (ni & 1) != half) continue;      // the other warp set drains this tile
            const Item w = get_item(g, it);
            if (do_stats && cur_q0 >= 0 && cur_q0 != w.q0) flush(cur_q0);
            cur_q0 = w.q0;
#ifdef B200SP_LEAN_TCG
            mbar_wait_sleep(&tfull[acc], tpar, g.epi_sleep);      // env B200SP_TCG_EPI_SLEEP: the 256 ns back-off returns after ~40 ns (r1j profile)
#else
            mbar_wait_sleep(&tfull[acc], tpar);
#endif
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
                if (ncol == 32) {
                    tc::tmem_ld32(t_row + c0, r);
                } else {
                    uint32_t r16[16];
                    tc::tmem_ld16(t_row + c0, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
                }
                tc::tmem_ld_wait();
                if (split_acc) {                // add the correction accumulator (fp32 round-to-nearest), 16 columns at a time
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (hh * 16 < ncol) {
                            uint32_t q16[16];
                            tc::tmem_ld16(t_row + g.BN + c0 + hh * 16, q16);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) r[hh * 16 + i] = __float_as_uint(__uint_as_float(r[hh * 16 + i]) + __uint_as_float(q16[i]));
                        }
                    }
                }
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

template <typename T, int ALAY, int BLAY, int EPI, int AMODE, int BMODE>
int launch_cfg(TcgArgs& a, cudaStream_t st) {
    using E = ET<T>;
    static bool attr_set = false;
    auto kern = tcgemm_kernel<T, ALAY, BLAY, EPI, AMODE, BMODE>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int numPt = ceil_div(a.P, BM);
    a.nkb = ceil_div(a.R, E::KE);
    // ---- partial last k-block of K-major operands ----
    {
        const int rem = a.R - (a.nkb - 1) * E::KE;             // 1..KE elements
        const int kstep = 2 * E::EPV;                          // elements per UMMA k-step (32 bytes)
        a.nks_last = ceil_div(rem, kstep);
        a.ac_last = 2 * a.nks_last;
    }
    // ---- tile width: the widest that fits shared memory; narrower while the grid does not cover the machine ----
    const int cap0 = E::TF32 ? 128 : 256;
    bool fits = false;
    for (int cap = cap0; cap >= 32 && !fits; cap -= (cap > 64 ? 32 : 16)) {
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
        const int atoms_b = ceil_div(BN * E::ES, 128);
        a.b_op_bytes = BLAY == TCG_LAY_KM ? BN * 128 : atoms_b * E::KE * 128;
        a.nb_pieces = a.b_op_bytes / 16;
        a.nb_slots = ceil_div(a.nb_pieces, PROD_T);
        const uint32_t bres_bytes = (uint32_t)a.nkb * E::NM * a.b_op_bytes;
        a.b_res = (EPI != TCG_EPI_ATOMIC && numQt == 1 && BMODE == XM_PLAIN && bres_bytes <= BRES_LIMIT) ? 1 : 0;
        a.op_stage_bytes = E::NM * (a.a_op_bytes + (a.b_res ? 0 : a.b_op_bytes));
        a.raw_stage_bytes = (APT + (AMODE == XM_DY ? APT : 0) + (a.b_res ? 0 : a.nb_slots)) * PROD_T * 16;
        const uint32_t fixed = EPI_W * 32 * STG_LD * 4 + EPI_W * 2 * BN * 4 + 256 + (a.b_res ? bres_bytes : 0);
        a.n_op = 2;
        a.n_raw = 0;
        for (int nr = MAX_RAW; nr >= 2; --nr) {
            if (a.n_op * a.op_stage_bytes + nr * a.raw_stage_bytes + fixed + 1088 <= SMEM_LIMIT) { a.n_raw = nr; break; }
        }
        if (a.n_raw == 0) continue;
        fits = true;
        while (a.n_op < MAX_OP && (a.n_op + 1) * a.op_stage_bytes + a.n_raw * a.raw_stage_bytes + fixed + 1088 <= SMEM_LIMIT) ++a.n_op;
        a.off_bres = a.n_op * a.op_stage_bytes;
        a.off_raw = a.off_bres + (a.b_res ? bres_bytes : 0);
        a.off_stg = a.off_raw + a.n_raw * a.raw_stage_bytes;
        a.off_stat = a.off_stg + EPI_W * 32 * STG_LD * 4;
        a.off_bar = (a.off_stat + EPI_W * 2 * BN * 4 + 15) & ~15u;
    }
    if (!fits) return B200SP_ENOSYS;
    const int numQt = a.numQt, BN = a.BN;
    // ---- reduction splits (wgrad only: results are accumulated atomically) ----
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
    const uint32_t smem = a.off_bar + 256 + 1024;
    {
        static int split_env = -1;
        if (split_env < 0) { const char* e = getenv("B200SP_TCG_SPLIT_ACC"); split_env = e ? atoi(e) : 1; }
        a.acc_cols = (E::TF32 && split_env && 2 * 2 * BN <= 512) ? 2 * BN : BN;
    }
    a.nacc = 4 * a.acc_cols <= 512 ? 4 : 2;
    a.split_epi = (BN <= 32 && numQt == 1 && EPI_SETS == 2) ? 1 : 0;
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.nacc * a.acc_cols)) cols <<= 1;
    a.tmem_cols = cols;
    const int total = numPt * numQt * a.splits;
    const int grid = total < NUM_SMS ? total : NUM_SMS;
    kern<<<grid, NT, smem, st>>>(a);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

template <typename T>
int launch_T(TcgArgs& a, const TcgProblem& p, cudaStream_t st) {
    const int am = p.a.mode, bm = p.b.mode;
    if (p.epi == TCG_EPI_FWD && p.a_lay == TCG_LAY_KM && p.b_lay == TCG_LAY_KM && bm == B200SP_VT_PLAIN) {
        if (am == B200SP_VT_PLAIN) return launch_cfg<T, TCG_LAY_KM, TCG_LAY_KM, TCG_EPI_FWD, XM_PLAIN, XM_PLAIN>(a, st);
        if (am == B200SP_VT_BNACT) return launch_cfg<T, TCG_LAY_KM, TCG_LAY_KM, TCG_EPI_FWD, XM_BNACT, XM_PLAIN>(a, st);
    }
    if (p.epi == TCG_EPI_DGRAD && p.a_lay == TCG_LAY_KM && p.b_lay == TCG_LAY_MM && bm == B200SP_VT_PLAIN) {
        if (am == B200SP_VT_PLAIN) return launch_cfg<T, TCG_LAY_KM, TCG_LAY_MM, TCG_EPI_DGRAD, XM_PLAIN, XM_PLAIN>(a, st);
        if (am == B200SP_VT_DY) return launch_cfg<T, TCG_LAY_KM, TCG_LAY_MM, TCG_EPI_DGRAD, XM_DY, XM_PLAIN>(a, st);
    }
    // split-K GEMMs with a plain fp32 red.add epilogue over K-major / mixed operands: the M <= 128 FC layers of the SPN
    // (one M tile: the K loop is the whole critical path, so it is split across CTAs)
    if (p.epi == TCG_EPI_ATOMIC && p.a_lay == TCG_LAY_KM && am == B200SP_VT_PLAIN && bm == B200SP_VT_PLAIN) {
        if (p.b_lay == TCG_LAY_KM) return launch_cfg<T, TCG_LAY_KM, TCG_LAY_KM, TCG_EPI_ATOMIC, XM_PLAIN, XM_PLAIN>(a, st);
        return launch_cfg<T, TCG_LAY_KM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_PLAIN, XM_PLAIN>(a, st);
    }
    if (p.epi == TCG_EPI_ATOMIC && p.a_lay == TCG_LAY_MM && p.b_lay == TCG_LAY_MM) {
        if (am == B200SP_VT_PLAIN && bm == B200SP_VT_PLAIN) return launch_cfg<T, TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_PLAIN, XM_PLAIN>(a, st);
        if (am == B200SP_VT_PLAIN && bm == B200SP_VT_BNACT) return launch_cfg<T, TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_PLAIN, XM_BNACT>(a, st);
        if (am == B200SP_VT_DY && bm == B200SP_VT_PLAIN) return launch_cfg<T, TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_DY, XM_PLAIN>(a, st);
        if (am == B200SP_VT_DY && bm == B200SP_VT_BNACT) return launch_cfg<T, TCG_LAY_MM, TCG_LAY_MM, TCG_EPI_ATOMIC, XM_DY, XM_BNACT>(a, st);
    }
    return B200SP_ENOSYS;
}

}  // namespace

int tcgemm_launch(const TcgProblem& p, cudaStream_t st) {
    const int epv = p.dtype == B200SP_F32 ? 4 : 8;
    // vector-piece granularity: every contiguous extent must be a whole number of 16-byte pieces
    if (p.Q % 4 || p.lda % epv || p.ldb % epv) return B200SP_ENOSYS;
    if (p.a_lay == TCG_LAY_KM ? (p.R % epv) : (p.P % epv)) return B200SP_ENOSYS;
    if (p.b_lay == TCG_LAY_KM ? (p.R % epv) : (p.Q % epv)) return B200SP_ENOSYS;
    if (p.b.mode == B200SP_VT_DY) return B200SP_ENOSYS;
    if (((uintptr_t)p.a.x | (uintptr_t)p.b.x | (uintptr_t)p.a.x2 | (uintptr_t)p.out) & 15) return B200SP_ENOSYS;
    TcgArgs a = {};
    a.a = p.a; a.b = p.b;
    a.P = p.P; a.Q = p.Q; a.R = p.R; a.lda = p.lda; a.ldb = p.ldb;
    a.ldo = p.ldo > 0 ? p.ldo : p.Q;
    if (a.ldo % 4) return B200SP_ENOSYS;
    a.a_dy = p.a.mode == B200SP_VT_DY;
    a.out = p.out; a.bias = p.bias; a.out_act = p.out_act;
    a.has_bnf = p.bnf != nullptr;
    if (p.bnf) a.bnf = *p.bnf;
    a.skip = p.skip; a.scale_out = p.scale_out;
    a.has_bnb = p.bnb != nullptr;
    if (p.bnb) a.bnb = *p.bnb;
    a.count = p.count;
    {
        static int wm = -1;
        if (wm < 0) { const char* e = getenv("B200SP_TCG_WAIT"); wm = e ? atoi(e) : 0; }
        a.wait_mode = wm;
        static int es = -1;
        if (es < 0) { const char* e = getenv("B200SP_TCG_EPI_SLEEP"); es = e ? atoi(e) : 256; }
        a.epi_sleep = (uint32_t)es;
    }
    if (p.dtype == B200SP_F32) return launch_T<float>(a, p, st);
    if (p.dtype == B200SP_BF16) return launch_T<bf16>(a, p, st);
    return B200SP_ENOSYS;
}

#ifdef TCG_TIMELINE
extern "C" int b200sp_tcg_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_tcg_tl, sizeof(long long) * 8 * 256);
}
#endif
