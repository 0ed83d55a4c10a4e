// gemm.cu -- 1x1-convolution GEMMs (forward, dgrad, wgrad) with fused BatchNorm/activation
// transforms on the operand loads and fused BatchNorm reductions in the epilogue.
//
// Generic problem:  C[P,Q] = sum_r Aop[p,r] * Bop[q,r]
//   forward : P=M  Q=N   R=K    A = f(X)[M,K]  (reduction contiguous, "KM")   B = W[N,K]   (KM)
//   dgrad   : P=M  Q=K   R=N    A = dY[M,N]    (KM)                           B = W[N,K]   (rows = reduction, "MM")
//   wgrad   : P=N  Q=K   R=M    A = dY[M,N]    (MM)                           B = f(X)[M,K] (MM), split over R
// In every case the channel index of the per-channel transform is the contiguous global column.
//
// Math: fp32 storage -> 3xTF32 split on mma.sync.m16n8k8 (error ~2^-21, indistinguishable from
// fp32 for the 1e-3 parity bar; SURVEY.md section 7 "Hard parts").  These layers are HBM-bound on
// B200, the tensor-bound layers go through the tcgen05 path (igemm_tc.cu).
#include <cstdlib>
#include "common.cuh"
#include "tcgemm.cuh"

int wgdirect_launch(const b200sp_vtensor* dy, const b200sp_vtensor* x, float* dw, int M, int N, int K, cudaStream_t st);   // wgdirect.cu
int pwdirect_fwd(const b200sp_vtensor* x, const float* w, const float* bias, int out_act, float* y, const b200sp_bnfwd* bn,
                 int M, int N, int K, cudaStream_t st);       // pwdirect.cu
int pwdirect_dgrad(const b200sp_vtensor* dy, const float* w, const float* skip, float scale_out, float* g, const b200sp_bnbwd* bn,
                   int M, int N, int K, cudaStream_t st);     // pwdirect.cu
extern "C" int b200sp_colsum_f32(const b200sp_vtensor* dy, float* out, int M, int N, int dtype, void* stream);

namespace {

constexpr int NT = 256;      // threads per CTA: 8 warps as 4 (M) x 2 (N)
constexpr int BK = 32;
enum { LAY_KM = 0, LAY_MM = 1 };
enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_ATOMIC = 2 };

struct GemmArgs {
    b200sp_vtensor a, b;
    int P, Q, R;
    int lda, ldb;
    int r_chunk;               // reduction range per blockIdx.z (multiple of BK)
    void* out;
    const float* bias;         // EPI_FWD
    int out_act;
    int has_bnf;
    b200sp_bnfwd bnf;
    const void* skip;          // EPI_DGRAD
    float scale_out;
    int has_bnb;               // mask / stats context present
    b200sp_bnbwd bnb;
    double count;
};

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
#ifdef B200SP_LEAN_TCG
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;      // == cvt.rna.tf32.f32 for finite v, 2 instead of 4 SASS instructions
#else
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
#endif
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Raw (untransformed) global tile held in registers while the previous tile is being multiplied.
template <int N4, bool TWO> struct TileRegs { float4 x[N4]; float4 x2[TWO ? N4 : 1]; };

template <typename T, int ROWS, int COLS, bool TWO>
__device__ __forceinline__ void tile_load(const b200sp_vtensor& t, int ld, int row0, int col0, int rowEnd, int colEnd,
                                          TileRegs<(ROWS * COLS / 4 + NT - 1) / NT, TWO>& r) {
    constexpr int C4 = COLS / 4, TOT = ROWS * C4, N4 = (TOT + NT - 1) / NT;
#pragma unroll
    for (int i = 0; i < N4; ++i) {
        int idx = threadIdx.x + i * NT;
        int rr = idx / C4, c4 = idx % C4;
        int gr = row0 + rr, gc = col0 + c4 * 4;
        bool ok = (TOT % NT == 0 || idx < TOT) && gr < rowEnd && gc < colEnd;
        float4 v = f4zero(), v2 = f4zero();
        if (ok) {
            size_t off = (size_t)gr * ld + gc;
            v = Vec4<T>::ld(reinterpret_cast<const T*>(t.x) + off);
            if (TWO) v2 = Vec4<T>::ld(reinterpret_cast<const T*>(t.x2) + off);
        }
        r.x[i] = v;
        if (TWO) r.x2[i] = v2;
    }
}
// transform + store to shared memory (a padded image of the global tile)
template <int ROWS, int COLS, int LDS, bool TWO>
__device__ __forceinline__ void tile_store(const b200sp_vtensor& t, int row0, int col0, int rowEnd, int colEnd,
                                           const TileRegs<(ROWS * COLS / 4 + NT - 1) / NT, TWO>& r, float* s) {
    constexpr int C4 = COLS / 4, TOT = ROWS * C4, N4 = (TOT + NT - 1) / NT;
#pragma unroll
    for (int i = 0; i < N4; ++i) {
        int idx = threadIdx.x + i * NT;
        if (TOT % NT != 0 && idx >= TOT) break;
        int rr = idx / C4, c4 = idx % C4;
        int gr = row0 + rr, gc = col0 + c4 * 4;
        float4 v = f4zero();
        if (gr < rowEnd && gc < colEnd) v = vt_apply4(t, r.x[i], TWO ? r.x2[i] : f4zero(), gc);
        *reinterpret_cast<float4*>(s + rr * LDS + c4 * 4) = v;
    }
}

template <typename T, int ALAY, int BLAY, int EPI, int MI, int NI>
__global__ void __launch_bounds__(NT, 2) gemm_kernel(const GemmArgs g) {
    constexpr int BM = 64 * MI, BN = 16 * NI;
    constexpr bool A2 = (EPI != EPI_FWD);            // dY operand carries (g, y)
    constexpr int A_ROWS = ALAY == LAY_KM ? BM : BK, A_COLS = ALAY == LAY_KM ? BK : BM;
    constexpr int B_ROWS = BLAY == LAY_KM ? BN : BK, B_COLS = BLAY == LAY_KM ? BK : BN;
    constexpr int LDA = A_COLS + (ALAY == LAY_KM ? 4 : 8);
    constexpr int LDB = B_COLS + (BLAY == LAY_KM ? 4 : 8);
    constexpr int A_N4 = (A_ROWS * A_COLS / 4 + NT - 1) / NT, B_N4 = (B_ROWS * B_COLS / 4 + NT - 1) / NT;
    __shared__ __align__(16) float As[A_ROWS * LDA];
    __shared__ __align__(16) float Bs[B_ROWS * LDB];
    __shared__ float s_red[2][4][BN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;          // "groupID" and "threadID_in_group" of the mma layout
    const int q0 = blockIdx.y * BN;
    const int r_begin = blockIdx.z * g.r_chunk;
    const int r_end = min(g.R, r_begin + g.r_chunk);
    const int numPt = (g.P + BM - 1) / BM;

    float csum[NI][2], csq[NI][2];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) { csum[ni][0] = csum[ni][1] = csq[ni][0] = csq[ni][1] = 0.f; }

    for (int pt = blockIdx.x; pt < numPt; pt += gridDim.x) {
        const int p0 = pt * BM;
        float acc[MI][NI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;

        TileRegs<A_N4, A2> ra;
        TileRegs<B_N4, false> rb;
        auto loadA = [&](int r0) {
            if (ALAY == LAY_KM) tile_load<T, A_ROWS, A_COLS, A2>(g.a, g.lda, p0, r0, g.P, r_end, ra);
            else                tile_load<T, A_ROWS, A_COLS, A2>(g.a, g.lda, r0, p0, r_end, g.P, ra);
        };
        auto loadB = [&](int r0) {
            if (BLAY == LAY_KM) tile_load<T, B_ROWS, B_COLS, false>(g.b, g.ldb, q0, r0, g.Q, r_end, rb);
            else                tile_load<T, B_ROWS, B_COLS, false>(g.b, g.ldb, r0, q0, r_end, g.Q, rb);
        };
        auto storeA = [&](int r0) {
            if (ALAY == LAY_KM) tile_store<A_ROWS, A_COLS, LDA, A2>(g.a, p0, r0, g.P, r_end, ra, As);
            else                tile_store<A_ROWS, A_COLS, LDA, A2>(g.a, r0, p0, r_end, g.P, ra, As);
        };
        auto storeB = [&](int r0) {
            if (BLAY == LAY_KM) tile_store<B_ROWS, B_COLS, LDB, false>(g.b, q0, r0, g.Q, r_end, rb, Bs);
            else                tile_store<B_ROWS, B_COLS, LDB, false>(g.b, r0, q0, r_end, g.Q, rb, Bs);
        };

        loadA(r_begin);
        loadB(r_begin);
        for (int r0 = r_begin; r0 < r_end; r0 += BK) {
            __syncthreads();                 // previous tile fully consumed
            storeA(r0);
            storeB(r0);
            __syncthreads();
            if (r0 + BK < r_end) { loadA(r0 + BK); loadB(r0 + BK); }
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                uint32_t ah[MI][4], al[MI][4], bh[NI][2], bl[NI][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int m = wm * (16 * MI) + mi * 16 + gq + (e & 1) * 8;
                        int k = ks * 8 + tq + (e >> 1) * 4;
                        float v = ALAY == LAY_KM ? As[m * LDA + k] : As[k * LDA + m];
                        split_tf32(v, ah[mi][e], al[mi][e]);
                    }
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        int n = wn * (8 * NI) + ni * 8 + gq;
                        int k = ks * 8 + tq + e * 4;
                        float v = BLAY == LAY_KM ? Bs[n * LDB + k] : Bs[k * LDB + n];
                        split_tf32(v, bh[ni][e], bl[ni][e]);
                    }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) {
                        mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                        mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                        mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                    }
            }
        }

        // ------------------------------ epilogue ------------------------------
        T* out = reinterpret_cast<T*>(g.out);
        const ActP oact = act_params(EPI == EPI_FWD ? g.out_act : g.bnb.act);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int col = q0 + wn * (8 * NI) + ni * 8 + 2 * tq;
            const bool cok = col < g.Q;           // Q is even, so col+1 is valid too
            float b0 = 0.f, b1 = 0.f, sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f, mu0 = 0.f, mu1 = 0.f, rs0 = 0.f, rs1 = 0.f;
            if (EPI == EPI_FWD && g.bias && cok) { b0 = g.bias[col]; b1 = g.bias[col + 1]; }
            if (EPI == EPI_DGRAD && g.has_bnb && cok) {
                if (g.bnb.scale) { sc0 = g.bnb.scale[col]; sc1 = g.bnb.scale[col + 1]; sh0 = g.bnb.shift[col]; sh1 = g.bnb.shift[col + 1]; }
                if (g.bnb.s1) { mu0 = g.bnb.mean[col]; mu1 = g.bnb.mean[col + 1]; rs0 = g.bnb.rstd[col]; rs1 = g.bnb.rstd[col + 1]; }
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = p0 + wm * (16 * MI) + mi * 16 + gq + h * 8;
                    const bool ok = cok && row < g.P;
                    float v0 = acc[mi][ni][h * 2], v1 = acc[mi][ni][h * 2 + 1];
                    const size_t off = (size_t)row * g.Q + col;
                    if (EPI == EPI_FWD) {
                        v0 = act_fwd(v0 + b0, oact);
                        v1 = act_fwd(v1 + b1, oact);
                        if (ok) {
                            Vec4<T>::st2(out + off, v0, v1);
                            if (g.has_bnf) { csum[ni][0] += v0; csum[ni][1] += v1; csq[ni][0] += v0 * v0; csq[ni][1] += v1 * v1; }
                        }
                    } else if (EPI == EPI_DGRAD) {
                        if (ok) {
                            v0 *= g.scale_out; v1 *= g.scale_out;
                            if (g.skip) {
                                const T* sk = reinterpret_cast<const T*>(g.skip) + off;
                                v0 += Vec4<T>::ld1(sk); v1 += Vec4<T>::ld1(sk + 1);
                            }
                            if (g.has_bnb) {
                                const T* yp = reinterpret_cast<const T*>(g.bnb.y) + off;
                                float y0 = Vec4<T>::ld1(yp), y1 = Vec4<T>::ld1(yp + 1);
                                v0 *= act_bwd(fmaf(y0, sc0, sh0), oact);
                                v1 *= act_bwd(fmaf(y1, sc1, sh1), oact);
                                if (g.bnb.s1) {
                                    csum[ni][0] += v0; csum[ni][1] += v1;
                                    csq[ni][0] += v0 * (y0 - mu0) * rs0; csq[ni][1] += v1 * (y1 - mu1) * rs1;
                                }
                            }
                            Vec4<T>::st2(out + off, v0, v1);
                        }
                    } else {
                        if (ok) {
                            float* o = reinterpret_cast<float*>(g.out) + off;      // even column, Q even: 8-byte aligned
                            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(v0), "f"(v1) : "memory");
                        }
                    }
                }
        }
    }

    // ---------------- fused BatchNorm reductions (forward stats / backward s1,s2) ----------------
    const bool stats = (EPI == EPI_FWD && g.has_bnf) || (EPI == EPI_DGRAD && g.has_bnb && g.bnb.s1);
    if (EPI != EPI_ATOMIC && stats) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float a = csum[ni][j], b = csq[ni][j];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (gq == 0) {
                    int c = wn * (8 * NI) + ni * 8 + 2 * tq + j;
                    s_red[0][wm][c] = a;
                    s_red[1][wm][c] = b;
                }
            }
        __syncthreads();
        if (tid < BN && q0 + tid < g.Q) {
            double a = (double)s_red[0][0][tid] + (double)s_red[0][1][tid] + (double)s_red[0][2][tid] + (double)s_red[0][3][tid];
            double b = (double)s_red[1][0][tid] + (double)s_red[1][1][tid] + (double)s_red[1][2][tid] + (double)s_red[1][3][tid];
            if (EPI == EPI_FWD) { atomicAdd(g.bnf.sum + q0 + tid, a); atomicAdd(g.bnf.sumsq + q0 + tid, b); }
            else                { atomicAdd(g.bnb.s1 + q0 + tid, a);  atomicAdd(g.bnb.s2 + q0 + tid, b); }
        }
        uint32_t* ticket = EPI == EPI_FWD ? g.bnf.ticket : g.bnb.ticket;
        if (grid_last_cta(ticket, gridDim.x * gridDim.y)) {
            for (int c = tid; c < g.Q; c += NT) {
                if (EPI == EPI_FWD) bn_fwd_finalize_channel(g.bnf, c, g.count);
                else                bn_bwd_finalize_channel(g.bnb, c, g.count);
            }
        }
    }
}

inline int tile_mi(int P) { return P <= 64 ? 1 : 2; }
inline int tile_ni(int Q) { return Q <= 16 ? 1 : (Q <= 32 ? 2 : 4); }

template <typename T, int ALAY, int BLAY, int EPI>
int launch_gemm(const GemmArgs& a, int splits, cudaStream_t st) {
    // tile choice: keep the padded-MMA waste low for the narrow MobileNetV2 layers
    const int mi = tile_mi(a.P), ni = tile_ni(a.Q);
    const int BM = 64 * mi, BN = 16 * ni;
    const int pt = ceil_div(a.P, BM), qt = ceil_div(a.Q, BN);
    dim3 grid(1, qt, splits);
    if (EPI == EPI_ATOMIC) grid.x = pt;
    else {
        int cap = max(1, (NUM_SMS * 4) / qt);
        grid.x = min(pt, cap);
    }
#define B200SP_GEMM_CASE(MI_, NI_) \
    if (mi == MI_ && ni == NI_) { gemm_kernel<T, ALAY, BLAY, EPI, MI_, NI_><<<grid, NT, 0, st>>>(a); B200SP_COUNT_LAUNCH(); B200SP_RETURN_LAST(); }
    B200SP_GEMM_CASE(1, 1) B200SP_GEMM_CASE(1, 2) B200SP_GEMM_CASE(1, 4)
    B200SP_GEMM_CASE(2, 1) B200SP_GEMM_CASE(2, 2) B200SP_GEMM_CASE(2, 4)
#undef B200SP_GEMM_CASE
    return B200SP_EINVAL;
}

inline b200sp_vtensor plain_vt(const void* p) {
    b200sp_vtensor t;
    t.x = p; t.x2 = nullptr; t.p0 = t.p1 = t.p2 = nullptr; t.mode = B200SP_VT_PLAIN; t.act = 0;
    return t;
}

// B200SP_GEMM=legacy forces the mma.sync kernels (A/B testing); default is the tcgen05 path with
// the mma.sync kernel as the fallback for shapes it does not cover (odd widths, unaligned pointers).
inline int gemm_mode() {      // 0 legacy only, 1 measured dispatch (default), 2 tcgen05 only
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_GEMM"); v = !e ? 1 : (e[0] == 'l' ? 0 : (e[0] == 't' ? 2 : 1)); }
    return v;
}
// Measured on B200 (profiles/r1_e_gemm_dispatch.txt, column "cuda-core"): for long-M layers with a small N x K this file's
// register-tiled mma.sync kernel (64*MI x 16*NI tiles, streaming split-M) beats the tcgen05 pipeline, whose 128-row tiles are
// mostly padding there and whose per-k-block hand-offs dominate.  op: 0 fwd, 1 dgrad, 2 wgrad.
inline bool prefer_cuda_cores(int op, int M, int N, int K) {
    if (M < 9408) return false;
    if (op == 2) return (long long)N * K <= 24576;
    static const int fwd_nk[][2] = {{16, 32}, {32, 144}, {32, 192}};
    static const int dgrad_nk[][2] = {{24, 144}, {32, 144}, {192, 32}, {64, 192}};
    if (op == 0) { for (auto& e : fwd_nk) if (e[0] == N && e[1] == K) return true; }
    if (op == 1) { for (auto& e : dgrad_nk) if (e[0] == N && e[1] == K) return true; }
    return false;
}
// second-generation tcgen05 kernel (tcgemm2.cu: TMA raw ring + lean converters) takes every shape it supports;
// B200SP_TCG2=0 restores the first-generation dispatch (tcgemm.cu + the measured mma.sync table) for A/B runs
inline bool tcg2_on() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_TCG2"); v = (e && e[0] == '0') ? 0 : 1; }      // default on (validated on B200, round 2)
    return v != 0;
}
inline bool tcg2_wgrad_on() {      // the weight-gradient form (both operands MN-major) can be switched separately: B200SP_TCG2_WGRAD=0
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200SP_TCG2_WGRAD"); v = (e && e[0] == '0') ? 0 : 1; }
    return tcg2_on() && v != 0;
}
inline bool use_tc(int op, int M, int N, int K) {
    const int m = gemm_mode();
    return m == 2 || (m == 1 && !prefer_cuda_cores(op, M, N, K));
}

}  // namespace

extern "C" int b200sp_pw_fwd(const b200sp_vtensor* x, const float* w, const float* bias, int out_act, void* y,
                             const b200sp_bnfwd* bn, int M, int N, int K, int dtype, void* stream) {
    if (!x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    if (dtype == B200SP_F32_TF32X1) {    // --use_fp16 mode: second-generation kernel only
        TcgProblem p = {};
        p.a = *x; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_KM;
        p.P = M; p.Q = N; p.R = K; p.lda = K; p.ldb = K; p.epi = TCG_EPI_FWD; p.dtype = dtype;
        p.out = y; p.bias = bias; p.out_act = out_act; p.bnf = bn; p.count = (double)M;
        const int rc = tcgemm2_launch(p, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
        dtype = B200SP_F32;              // shapes the kernel does not take run at full precision
    }
    if (dtype == B200SP_F32) {           // round-2 candidate (B200SP_PWDIRECT=1): exact-fp32 FFMA kernel for the long-M / tiny-N*K shapes
        const int rc = pwdirect_fwd(x, w, bias, out_act, (float*)y, bn, M, N, K, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype == B200SP_BF16 || tcg2_on() || use_tc(0, M, N, K)) {
        TcgProblem p = {};
        p.a = *x; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_KM;
        p.P = M; p.Q = N; p.R = K; p.lda = K; p.ldb = K; p.epi = TCG_EPI_FWD; p.dtype = dtype;
        p.out = y; p.bias = bias; p.out_act = out_act; p.bnf = bn; p.count = (double)M;
        int rc = tcg2_on() ? tcgemm2_launch(p, (cudaStream_t)stream) : B200SP_ENOSYS;
        if (rc == B200SP_ENOSYS && (dtype == B200SP_BF16 || use_tc(0, M, N, K))) rc = tcgemm_launch(p, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    if (N % 2 || K % 4) return B200SP_EINVAL;
    GemmArgs a = {};
    a.a = *x; a.b = plain_vt(w);
    a.P = M; a.Q = N; a.R = K; a.lda = K; a.ldb = K; a.r_chunk = ((K + BK - 1) / BK) * BK;
    a.out = y; a.bias = bias; a.out_act = out_act;
    a.has_bnf = bn != nullptr;
    if (bn) a.bnf = *bn;
    a.count = (double)M;
    return launch_gemm<float, LAY_KM, LAY_KM, EPI_FWD>(a, 1, (cudaStream_t)stream);
}

extern "C" int b200sp_pw_dgrad(const b200sp_vtensor* dy, const float* w, const void* skip, float scale_out, void* g,
                               const b200sp_bnbwd* bn, int M, int N, int K, int dtype, void* stream) {
    if (!dy) return B200SP_EINVAL;
    if (dtype == B200SP_F32_TF32X1) {
        TcgProblem p = {};
        p.a = *dy; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_MM;
        p.P = M; p.Q = K; p.R = N; p.lda = N; p.ldb = K; p.epi = TCG_EPI_DGRAD; p.dtype = dtype;
        p.out = g; p.skip = skip; p.scale_out = scale_out; p.bnb = bn; p.count = (double)M;
        const int rc = tcgemm2_launch(p, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
        dtype = B200SP_F32;
    }
    if (dtype == B200SP_F32) {           // round-2 candidate (B200SP_PWDIRECT=1)
        const int rc = pwdirect_dgrad(dy, w, (const float*)skip, scale_out, (float*)g, bn, M, N, K, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype == B200SP_BF16 || tcg2_on() || use_tc(1, M, N, K)) {
        TcgProblem p = {};
        p.a = *dy; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_MM;
        p.P = M; p.Q = K; p.R = N; p.lda = N; p.ldb = K; p.epi = TCG_EPI_DGRAD; p.dtype = dtype;
        p.out = g; p.skip = skip; p.scale_out = scale_out; p.bnb = bn; p.count = (double)M;
        int rc = tcg2_on() ? tcgemm2_launch(p, (cudaStream_t)stream) : B200SP_ENOSYS;
        if (rc == B200SP_ENOSYS && (dtype == B200SP_BF16 || use_tc(1, M, N, K))) rc = tcgemm_launch(p, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    if (N % 4 || K % 4) return B200SP_EINVAL;
    GemmArgs a = {};
    a.a = *dy;
    if (a.a.mode != B200SP_VT_DY) {   // plain gradient: express as DY with unit coefficients is wasteful -> alias x2 = x, handled by mode
        a.a.x2 = a.a.x;
    }
    a.b = plain_vt(w);
    a.P = M; a.Q = K; a.R = N; a.lda = N; a.ldb = K; a.r_chunk = ((N + BK - 1) / BK) * BK;
    a.out = g; a.skip = skip; a.scale_out = scale_out;
    a.has_bnb = bn != nullptr;
    if (bn) a.bnb = *bn;
    a.count = (double)M;
    return launch_gemm<float, LAY_KM, LAY_MM, EPI_DGRAD>(a, 1, (cudaStream_t)stream);
}

extern "C" int b200sp_pw_wgrad(const b200sp_vtensor* dy, const b200sp_vtensor* x, float* dw, float* dbias,
                               int M, int N, int K, int dtype, void* stream) {
    if (!dy || !x || x->mode == B200SP_VT_DY) return B200SP_EINVAL;
    if (dtype == B200SP_F32 || dtype == B200SP_F32_TF32X1) {      // long-M / tiny-N*K layers: streaming fp32 reduction (wgdirect.cu)
        const int rc = wgdirect_launch(dy, x, dw, M, N, K, (cudaStream_t)stream);
        if (rc == 0 && dbias) return b200sp_colsum_f32(dy, dbias, M, N, B200SP_F32, stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype == B200SP_F32_TF32X1) {
        TcgProblem p = {};
        p.a = *dy; p.b = *x; p.a_lay = TCG_LAY_MM; p.b_lay = TCG_LAY_MM;
        p.P = N; p.Q = K; p.R = M; p.lda = N; p.ldb = K; p.epi = TCG_EPI_ATOMIC; p.dtype = dtype;
        p.out = dw;
        const int rc = tcgemm2_launch(p, (cudaStream_t)stream);
        if (rc == 0 && dbias) return b200sp_colsum_f32(dy, dbias, M, N, B200SP_F32, stream);
        if (rc != B200SP_ENOSYS) return rc;
        dtype = B200SP_F32;
    }
    if (dtype == B200SP_BF16 || tcg2_on() || use_tc(2, M, N, K)) {
        TcgProblem p = {};
        p.a = *dy; p.b = *x; p.a_lay = TCG_LAY_MM; p.b_lay = TCG_LAY_MM;
        p.P = N; p.Q = K; p.R = M; p.lda = N; p.ldb = K; p.epi = TCG_EPI_ATOMIC; p.dtype = dtype;
        p.out = dw;
        int rc = tcg2_wgrad_on() ? tcgemm2_launch(p, (cudaStream_t)stream) : B200SP_ENOSYS;
        if (rc == B200SP_ENOSYS && (dtype == B200SP_BF16 || use_tc(2, M, N, K))) rc = tcgemm_launch(p, (cudaStream_t)stream);
        if (rc == 0 && dbias) return b200sp_colsum_f32(dy, dbias, M, N, dtype, stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    if (dtype != B200SP_F32) return B200SP_ENOSYS;
    if (N % 4 || K % 4) return B200SP_EINVAL;
    GemmArgs a = {};
    a.a = *dy;
    if (a.a.mode != B200SP_VT_DY) a.a.x2 = a.a.x;
    a.b = *x;
    a.P = N; a.Q = K; a.R = M; a.lda = N; a.ldb = K;
    const int mi = tile_mi(N), ni = tile_ni(K);
    const int tiles = ceil_div(N, 64 * mi) * ceil_div(K, 16 * ni);
    int splits = max(1, min((NUM_SMS * 4) / tiles, ceil_div(M, BK * 4)));
    int chunk = ceil_div(ceil_div(M, splits), BK) * BK;
    splits = ceil_div(M, chunk);
    a.r_chunk = chunk;
    a.out = dw;
    int rc = launch_gemm<float, LAY_MM, LAY_MM, EPI_ATOMIC>(a, splits, (cudaStream_t)stream);
    if (rc) return rc;
    if (dbias) return b200sp_colsum_f32(dy, dbias, M, N, dtype, stream);
    return 0;
}

// ---- strided variants (grouped convolutions as per-group GEMMs over column slices; SPN, src/nets/spn.py:65,73,76) ----
// tensor-core path only: every extent must satisfy the 16-byte granularity rules of tcgemm_launch.
extern "C" int b200sp_gemm_fwd(const b200sp_vtensor* x, int ldx, const float* w, const float* bias, int out_act, void* y, int ldy,
                               int M, int N, int K, int dtype, void* stream) {
    if (!x || x->mode == B200SP_VT_DY || dtype != B200SP_F32) return B200SP_EINVAL;
    TcgProblem p = {};
    p.a = *x; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_KM;
    p.P = M; p.Q = N; p.R = K; p.lda = ldx; p.ldb = K; p.ldo = ldy; p.epi = TCG_EPI_FWD; p.dtype = dtype;
    p.out = y; p.bias = bias; p.out_act = out_act; p.bnf = nullptr; p.count = (double)M;
    return tcgemm_launch(p, (cudaStream_t)stream);
}

extern "C" int b200sp_gemm_dgrad(const b200sp_vtensor* dy, int lddy, const float* w, const void* skip, float scale_out, void* g,
                                 const b200sp_bnbwd* bn, int M, int N, int K, int dtype, void* stream) {
    if (!dy || dtype != B200SP_F32) return B200SP_EINVAL;
    TcgProblem p = {};
    p.a = *dy; p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_MM;
    p.P = M; p.Q = K; p.R = N; p.lda = lddy; p.ldb = K; p.epi = TCG_EPI_DGRAD; p.dtype = dtype;
    p.out = g; p.skip = skip; p.scale_out = scale_out; p.bnb = bn; p.count = (double)M;
    return tcgemm_launch(p, (cudaStream_t)stream);
}

extern "C" int b200sp_gemm_wgrad(const b200sp_vtensor* dy, int lddy, const b200sp_vtensor* x, int ldx, float* dw,
                                 int M, int N, int K, int dtype, void* stream) {
    if (!dy || !x || x->mode == B200SP_VT_DY || dtype != B200SP_F32) return B200SP_EINVAL;
    TcgProblem p = {};
    p.a = *dy; p.b = *x; p.a_lay = TCG_LAY_MM; p.b_lay = TCG_LAY_MM;
    p.P = N; p.Q = K; p.R = M; p.lda = lddy; p.ldb = ldx; p.epi = TCG_EPI_ATOMIC; p.dtype = dtype;
    p.out = dw;
    return tcgemm_launch(p, (cudaStream_t)stream);
}

// ---- split-K FC layers (M <= 128 rows: SPN fc6-fc11, src/nets/spn.py:80-99) ------------------------------------------
// With a single 128-row M tile the K loop (K/8 x 3 tcgen05.mma, ~115 cycles each) is the whole critical path of a CTA;
// splitting it across CTAs and reducing with fp32 red.add turns 267 us into ~50 us for fc6.  The output must be
// zero (fwd) or hold the value to accumulate onto (dgrad: the other branch's gradient) on entry.
int fc_fwd_stream(const float* x, const float* w, float* y_acc, int M, int N, int K, cudaStream_t st);         // fcstream.cu
int fc_dgrad_stream(const float* dy, const float* w, float* dx_acc, int M, int N, int K, cudaStream_t st);     // fcstream.cu
extern "C" int b200sp_fc_fwd_splitk(const float* x, const float* w, float* y_acc, int M, int N, int K, void* stream) {
    {   // batches up to 32 rows: the layer is a pure weight stream (fcstream.cu); larger ones take the split-K tensor-core GEMM
        const int rc = fc_fwd_stream(x, w, y_acc, M, N, K, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    TcgProblem p = {};
    p.a = plain_vt(x); p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_KM;
    p.P = M; p.Q = N; p.R = K; p.lda = K; p.ldb = K; p.epi = TCG_EPI_ATOMIC; p.dtype = B200SP_F32;
    p.out = y_acc;
    return tcgemm_launch(p, (cudaStream_t)stream);
}
extern "C" int b200sp_fc_dgrad_splitk(const float* dy, const float* w, float* dx_acc, int M, int N, int K, void* stream) {
    {
        const int rc = fc_dgrad_stream(dy, w, dx_acc, M, N, K, (cudaStream_t)stream);
        if (rc != B200SP_ENOSYS) return rc;
    }
    TcgProblem p = {};
    p.a = plain_vt(dy); p.b = plain_vt(w); p.a_lay = TCG_LAY_KM; p.b_lay = TCG_LAY_MM;
    p.P = M; p.Q = K; p.R = N; p.lda = N; p.ldb = K; p.epi = TCG_EPI_ATOMIC; p.dtype = B200SP_F32;
    p.out = dx_acc;
    return tcgemm_launch(p, (cudaStream_t)stream);
}
