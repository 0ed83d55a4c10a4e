// tcgemm.cuh -- internal interface between the C-ABI GEMM entry points (gemm.cu) and the
// tcgen05 kernel (tcgemm.cu).
#pragma once
#include "common.cuh"

enum { TCG_LAY_KM = 0, TCG_LAY_MM = 1 };
enum { TCG_EPI_FWD = 0, TCG_EPI_DGRAD = 1, TCG_EPI_ATOMIC = 2 };

// C[P,Q] = sum_r A(p,r) * B(q,r).  Layout KM: the operand is stored [MN][R] (reduction contiguous),
// MM: stored [R][MN].  The per-channel transform of a virtual tensor indexes the CONTIGUOUS dimension.
struct TcgProblem {
    b200sp_vtensor a, b;
    int a_lay, b_lay;
    int P, Q, R;
    int lda, ldb;
    int ldo;                   // row stride of `out` (FWD / DGRAD); 0 => Q
    int epi;
    int dtype;                 // B200SP_F32 (3xTF32 math) | B200SP_BF16
    void* out;                 // [P,Q] T (FWD, DGRAD) or float (ATOMIC, accumulated)
    // FWD
    const float* bias;
    int out_act;
    const b200sp_bnfwd* bnf;
    // DGRAD
    const void* skip;
    float scale_out;
    const b200sp_bnbwd* bnb;
    double count;
};

// returns 0 on success, B200SP_ENOSYS when the shape is outside what the tensor-core path supports
// (the caller then uses the CUDA-core/mma.sync fallback kernel), or a cudaError_t.
int tcgemm_launch(const TcgProblem& p, cudaStream_t st);

// second-generation kernel (tcgemm2.cu): raw operands staged by TMA, lean smem->smem converters; same contract.
int tcgemm2_launch(const TcgProblem& p, cudaStream_t st);

// ---- presplit path (tcgemm2.cu PRE mode + opsplit.cu) ------------------------------------------------------------------------
// One operand of a GEMM: the virtual tensor stored [rows][ld] with `cols` channels (the per-channel transform indexes the column),
// written as nm planes (tf32 hi [, lo]) of out[.][out_ld], either in the same orientation or transposed ([cols][rows]).
struct OpSplitJob {
    b200sp_vtensor t;
    int rows, cols, ld;
    float* out;
    int out_ld;
    size_t plane;              // floats between the hi and the lo plane
    int trans, nm;
};
int opsplit_launch(const OpSplitJob& a, const OpSplitJob& b, cudaStream_t st);
// library-owned scratch for the presplit operands (b200sp_set_workspace); {nullptr, 0} until the host provides one
void tcg_workspace(float** base, size_t* bytes);
// 0 | B200SP_ENOSYS (shape not worth it / no workspace: the caller continues with tcgemm2_launch) | cudaError_t
int tcgemm2_presplit(const TcgProblem& p, cudaStream_t st);
