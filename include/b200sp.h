/*
 * b200sp.h -- C-ABI of libb200sp.so: the sm_100a kernels behind the
 * speedplusbaseline CNN-training hot path (KRN / SPN / style-aug / DANN / AdamW).
 *
 * The reference (pure Python) has no FFI of its own: every entry point below
 * replaces one torch op *call site* of the reference, cited per function as
 * /root/reference/<file>:<line>.  The binding a maintainer adds is a ctypes
 * stub; speedplusbaseline_b200/_lib.py is that stub.
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless the name ends in _host;
 *  - activations are NHWC, contiguous, element type selected by `dtype`
 *    (B200SP_F32 = float, B200SP_BF16 = __nv_bfloat16); parameters, BatchNorm
 *    statistics and optimizer state are always float;
 *  - every launcher enqueues on `stream` (a cudaStream_t passed as void*),
 *    never synchronises, never allocates, and returns 0 or a cudaError_t /
 *    negative B200SP_E* code;
 *  - a "virtual tensor" (b200sp_vtensor) is a raw conv output plus the
 *    per-channel affine + activation that the reference would have
 *    materialised with batch_norm/relu6 -- consumers apply it on load, so the
 *    normalised tensor never round-trips through HBM.
 */
#ifndef B200SP_H
#define B200SP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SP_F32 0
#define B200SP_BF16 1
/* accepted by the b200sp_pw_* GEMM entry points only: fp32 storage, SINGLE-pass TF32 tensor-core math (operands rounded to a
 * 10-bit mantissa -- fp16's -- with fp32 range and fp32 accumulation).  This is what `--use_fp16` selects (config.py:39,
 * trainer.py:73-94): the autocast numerics of the GEMMs without a loss scaler, because the exponent range stays fp32's. */
#define B200SP_F32_TF32X1 2

#define B200SP_ACT_NONE 0
#define B200SP_ACT_RELU 1
#define B200SP_ACT_RELU6 2
#define B200SP_ACT_LEAKY02 3   /* LeakyReLU(0.2), park2019.py:66 */
#define B200SP_ACT_SIGMOID 4   /* forward only, ghiasi.py:135 */

#define B200SP_VT_PLAIN 0      /* v = x                                   */
#define B200SP_VT_BNACT 1      /* v = act(x * p0[c] + p1[c])              */
#define B200SP_VT_DY 2         /* v = p0[c]*x + p1[c]*x2 + p2[c]  (BN backward folded into the load) */

#define B200SP_EINVAL (-22)
#define B200SP_ENOSYS (-38)

typedef struct b200sp_vtensor {
    const void *x;      /* raw tensor */
    const void *x2;     /* second tensor for B200SP_VT_DY (the saved raw conv output), else NULL */
    const float *p0;    /* per-channel parameters, see modes above */
    const float *p1;
    const float *p2;
    int32_t mode;
    int32_t act;
} b200sp_vtensor;

/* Training-mode BatchNorm statistics fused into the producing conv's epilogue
 * (replaces the stats half of torch batch_norm at park2019.py:48,52,64 and
 * torchvision mobilenetv2.py:38,52).  The last CTA to finish turns the
 * per-channel sums into scale/shift (+ saved mean/rstd), updates the running
 * statistics (momentum, unbiased variance) and re-zeroes sum/sumsq/ticket. */
typedef struct b200sp_bnfwd {
    double *sum;              /* [C] zero on entry, zero on exit */
    double *sumsq;            /* [C] */
    uint32_t *ticket;         /* [1] zero on entry, zero on exit */
    const float *gamma;       /* [C] */
    const float *beta;        /* [C] */
    float *running_mean;      /* [C] updated in place (may be NULL) */
    float *running_var;       /* [C] */
    float *scale;             /* [C] out: gamma * rstd */
    float *shift;             /* [C] out: beta - mean * scale */
    float *mean;              /* [C] out */
    float *rstd;              /* [C] out */
    float momentum;
    float eps;
} b200sp_bnfwd;

/* BatchNorm backward reductions fused into the epilogue of the kernel that
 * produces g = dL/d(act output) * act'(z):  s1 = sum g, s2 = sum g*xhat.
 * The last CTA writes dgamma/dbeta (accumulating) and the coefficients of
 * dy = cA*g + cB*y + cC, which the next consumer applies on load (VT_DY). */
typedef struct b200sp_bnbwd {
    double *s1;               /* [C] zero on entry/exit */
    double *s2;               /* [C] */
    uint32_t *ticket;
    const void *y;            /* saved raw conv output of the BN input (same shape as g) */
    const float *scale;       /* [C] forward-saved */
    const float *shift;
    const float *mean;
    const float *rstd;
    float *cA;                /* [C] out */
    float *cB;
    float *cC;
    float *dgamma;            /* [C] += s2 */
    float *dbeta;             /* [C] += s1 */
    int32_t act;              /* activation that followed this BN */
    int32_t pad_;
} b200sp_bnbwd;

/* ---- library ------------------------------------------------------------ */
int b200sp_version(void);
/* number of kernel launches issued through this library since load (the bench's gpu_launches) */
int64_t b200sp_launch_count(void);

/* Scratch for the presplit GEMM route (tcgemm2.cu PRE mode + opsplit.cu): the 1x1 convolutions whose activation operand has at
 * most 4096 rows (the 7x7 layers of the KRN at batch 48) write their operands once as tf32 hi/lo planes into this buffer and run a
 * converter-free TMA -> tcgen05 kernel on them.  `base`: device memory, 256-byte aligned, owned by the caller and alive until
 * replaced (NULL withdraws it); 96 MB covers every layer of the KRN.  Without a workspace the general kernel is used.  The
 * first half is reused by every eligible forward / data-gradient call, the second half by every eligible weight-gradient
 * call: issue each kind on ONE stream at a time (the engines run the weight gradients on a side stream).  The pointer is
 * process-global: one device per process (the launch model of this library: one process per GPU). */
int b200sp_set_workspace(void *base, size_t bytes);

/* tcgen05 plumbing self-test (tc_probe.cu): D[128,N] = A * B^T on one CTA with operands staged in the
 * library's swizzled shared-memory formats.  mode 0 tf32, 1 3xTF32, 2 bf16; *_major 0: operand given
 * as [MN][Ktot] (reduction contiguous), 1: as [Ktot][MN].  Ktot = nkb * (32 fp32 | 64 bf16). */
int b200sp_tc_probe(const void *A, const void *B, float *D, int N, int nkb, int mode,
                    int a_major, int b_major, int variant, void *stream);
/* tcgen05.mma issue-rate probe (mma_probe.cu; measurement tool): `grid` CTAs each issue `iters` back-to-back MMAs (M 128,
 * K 32 bytes, width N) and write the elapsed clock64 cycles to out_cycles[cta].  kind 0 tf32 | 1 bf16; a_src 0 shared memory
 * | 1 TMEM; layout 0 SWIZZLE_128B | 1 none | 2 SWIZZLE_64B | 3 SWIZZLE_32B; rotate 1 walks k-steps / stages like a main loop */
int b200sp_mma_probe(int kind, int a_src, int layout, int N, int iters, int rotate, int grid, long long *out_cycles, void *stream);

/* ---- convolutions -------------------------------------------------------- */
/* 3x3 stride-2 pad-1 stem, 3 -> Cout(32), NCHW float input (what the loader yields),
 * NHWC output.  torchvision mobilenetv2.py:126 via park2019.py:107-108. */
int b200sp_stem_fwd(const float *x_nchw, const float *w /*[32,3,3,3]*/, void *y,
                    const b200sp_bnfwd *bn /*NULL in eval*/, int B, int H, int W, int dtype, void *stream);
int b200sp_stem_wgrad(const float *x_nchw, const b200sp_vtensor *dy, float *dw /*[32,27] +=*/,
                      int B, int H, int W, int dtype, void *stream);

/* 1x1 convolution as GEMM  Y[M,N] = f(X)[M,K] * W[N,K]^T   (park2019.py:51,64;
 * torchvision mobilenetv2.py:38,52).  bias may be NULL; out_act applied after bias
 * (domain head revgrad.py:76-77, SPN FC spn.py:80-99). */
int b200sp_pw_fwd(const b200sp_vtensor *x, const float *w, const float *bias, int out_act, void *y,
                  const b200sp_bnfwd *bn /*may be NULL*/, int M, int N, int K, int dtype, void *stream);
/* dX[M,K] = dY[M,N] * W[N,K] (+ skip);  then g = dX * act'(z) and the BN-backward
 * reductions of the *input's* BatchNorm when `bn` is given.  scale_out multiplies the
 * result first (gradient reversal: -alpha, revgrad.py:52-56). */
int b200sp_pw_dgrad(const b200sp_vtensor *dy, const float *w, const void *skip, float scale_out,
                    void *g, const b200sp_bnbwd *bn /*may be NULL*/, int M, int N, int K, int dtype, void *stream);
/* dW[N,K] += dY[M,N]^T * f(X)[M,K];  dbias[N] += column sums of dY when dbias != NULL */
int b200sp_pw_wgrad(const b200sp_vtensor *dy, const b200sp_vtensor *x, float *dw, float *dbias,
                    int M, int N, int K, int dtype, void *stream);

/* depthwise 3x3 pad 1, stride 1|2, weights [9][C] (tap-major).  park2019.py:47;
 * torchvision mobilenetv2.py:45-49. */
int b200sp_dw_fwd(const b200sp_vtensor *x, const float *w9c, void *y, const b200sp_bnfwd *bn,
                  int B, int H, int W, int C, int stride, int dtype, void *stream);
/* fused depthwise dgrad + wgrad + activation/BN backward of the conv INPUT:
 *   g_in = (dgrad(dy) [+ skip]) * act'(z_in);  dw9c += wgrad;  BN reductions for the input's BN. */
int b200sp_dw_bwd(const b200sp_vtensor *dy, const b200sp_vtensor *x, const float *w9c, const void *skip,
                  void *g_in, float *dw9c, const b200sp_bnbwd *bn,
                  int B, int H, int W, int C, int stride, int dtype, void *stream);

/* ---- normalisation / elementwise ----------------------------------------- */
/* out = y*scale + shift (+ residual), materialising a block output (mobilenetv2.py:61-62) */
int b200sp_bn_apply(const void *y, const float *scale, const float *shift, const void *residual,
                    int act, void *out, int64_t M, int C, int dtype, void *stream);
/* eval mode: scale/shift from running statistics for `n` channels at once */
int b200sp_bn_eval_affine(const float *gamma, const float *beta, const float *rmean, const float *rvar,
                          float eps, float *scale, float *shift, int64_t n, void *stream);
/* standalone versions of the fused "last CTA" steps */
int b200sp_bn_fwd_finalize(const b200sp_bnfwd *bn, int C, double count, void *stream);
int b200sp_bn_bwd_finalize(const b200sp_bnbwd *bn, int C, double count, void *stream);
/* s1/s2 reductions for a materialised gradient (g, y) pair, then finalize */
int b200sp_bn_bwd_reduce(const void *g, const b200sp_bnbwd *bn, int64_t M, int C, int dtype, void *stream);
int b200sp_add_i64(int64_t *p, int64_t n, int64_t v, void *stream);     /* num_batches_tracked += 1 */

/* RouterV2 space-to-depth + concat (park2019.py:70-80): out[B,h,w,4*Cr + C1] */
int b200sp_reorg_cat_fwd(const b200sp_vtensor *xr /*[B,2h,2w,Cr]*/, const b200sp_vtensor *x1 /*[B,h,w,C1]*/,
                         void *out, int B, int h, int w, int Cr, int C1, int dtype, void *stream);
int b200sp_reorg_cat_bwd(const void *dcat, void *g_r, void *g_1, const b200sp_bnbwd *bn_r,
                         const b200sp_bnbwd *bn_1, int B, int h, int w, int Cr, int C1, int dtype, void *stream);

/* ---- KRN head + loss ------------------------------------------------------ */
/* 7x7 valid conv 1024 -> 22 == FC over the NHWC-flattened map (park2019.py:121,139).
 * logits[B,N] must hold the bias on entry (b200sp_krn_loss_prep); accumulates atomically. */
int b200sp_head_fwd(const b200sp_vtensor *x, const float *w /*[N][HW*C]*/, float *logits,
                    int B, int HWC, int C, int N, int dtype, void *stream);
int b200sp_head_bias(const float *bias, float *logits, int B, int N, void *stream);
/* park2019.py:142-160: loss3 = {loss, loss_x, loss_y}; dlogits = 2(x-t)/B * loss_scale[0] */
int b200sp_krn_loss(const float *logits, const float *target /*[B,2,N/2]*/, float *loss3, float *dlogits,
                    float *dbias /* += */, const float *loss_scale /*NULL => 1*/, int B, int N, void *stream);
int b200sp_head_bwd(const float *dlogits, const b200sp_vtensor *x, const float *w, void *g, float *dw /* += */,
                    float *dbias /* += */, const b200sp_bnbwd *bn, int B, int HWC, int C, int N, int dtype, void *stream);

/* ---- style augmentation (src/styleaug/ghiasi.py:6-135, styleAugmentor.py:44-68), forward only -------------- */
/* Dense k x k convolution as a shifted GEMM on tcgen05 with TMA-staged bf16 operands (csrc/convtc.cu).
 * Replaces ReflectionPad2d + Conv2d (ghiasi.py:11-12,36-37,73-81) and Upsample + ... (ghiasi.py:34-37): padding,
 * stride-2 phase split and nearest upsampling are materialised by whoever wrote the input planes
 * (b200sp_in_apply / b200sp_sa_prep).  planes[p]: bf16 [B*Hq*Wq (+8 rows of slack)][C]; chunk j contributes
 * plane rows m + shift_j, 128 bytes starting at channel c0_j, against weight columns [64j, 64j+64) of
 * w = bf16 [N_pad][64*n_chunks].  out: fp32 [B][Ho][Wo][N_out]; stats (may be NULL): fp32 [B][2][N_pad]
 * accumulated (+=) per-(image, channel) sum and sum of squares over the Ho*Wo valid outputs. */
#define B200SP_CONVTC_MAX_CHUNKS 64
typedef struct b200sp_convtc_chunk { int32_t plane, c0, shift, pad_; } b200sp_convtc_chunk;
typedef struct b200sp_convtc_desc {
    const void *planes[4];
    const void *w;
    float *out;
    float *stats;
    int32_t C, B, Hq, Wq, Ho, Wo, N_pad, N_out, n_chunks, pad_;
    b200sp_convtc_chunk chunks[B200SP_CONVTC_MAX_CHUNKS];
    /* output scatter (all zero => dense [B][Ho][Wo]): valid grid position (ph, pw) of image b is stored at pixel
     * (ph*sy + oy, pw*sx + ox) of an [B][OH][OW][N_out] tensor -- the four phase convolutions of a sub-pixel
     * (upsample x2 + 3x3) layer interleave their outputs this way */
    int32_t OH, OW, sy, sx, oy, ox;
} b200sp_convtc_desc;
int b200sp_convtc_fwd(const b200sp_convtc_desc *d, void *stream);

/* second half of a row-decomposed k x k convolution with Co <= 4 outputs (ghiasi.py:121): T [B][Ho][Wq][Nt] holds, per
 * plane-grid pixel, the k*Co partial sums over (kh, c) computed by b200sp_convtc_fwd with k row-shift taps;
 * out[b,h,w,co] = sum_kw T[b,h,w+kw][kw*Co+co] (fp32 [B][Ho][Wo][N_out]) and stats [B][2][N_pad] += per-image sum / sum^2 */
int b200sp_conv_kwsum(const float *T, float *out, float *stats, int B, int Ho, int Wo, int Wq, int Nt, int k, int Co, int N_out,
                      int N_pad, void *stream);
/* NCHW fp32 image [B,3,H,W] -> reflection-padded NHWC bf16 plane [B][H+2p][W+2p][Cd] (channels >= 3 zero) */
int b200sp_sa_prep(const float *x_nchw, void *plane, int B, int H, int W, int pad, int Cd, void *stream);
/* InstanceNorm2d(affine=False, eps) statistics -> per-(image, channel) scale/shift, folding the conditional
 * gamma/beta (ghiasi.py:57-61,94-103; NULL => 1/0):  scale = gamma*rstd, shift = beta - mean*scale.
 * gamma/beta: [B][gb_stride] rows.  Re-zeroes stats. */
int b200sp_in_finalize(float *stats, const float *gamma, const float *beta, int gb_stride, float *scale, float *shift,
                       int B, int C, int N_pad, int HW, float eps, void *stream);
/* v = act(raw*scale + shift) (+ res_in), written (a) as the NEXT conv's bf16 input plane(s): reflection padding
 * `pad`, nearest upsampling `up` (1|2), stride-2 phase split `ps` (1|2 -> 1|4 planes of [B][Hd][Wd][Cd]);
 * (b) optionally as the fp32 residual stream res_out [B][Hs][Ws][C] (interior pixels, ps == up == 1 only).
 * pad_mode 0: reflection (ReflectionPad2d); 1: edge replication (what reflection of a x2-nearest-upsampled image is in
 * source coordinates: used by the sub-pixel form of the upsampling convolutions). */
typedef struct b200sp_in_apply_desc {
    const float *raw;        /* [B][Hs][Ws][Cs] */
    const float *scale, *shift;   /* [B][C] */
    const float *res_in;     /* fp32 [B][Hs][Ws][C] or NULL */
    float *res_out;          /* or NULL */
    void *planes[4];
    int32_t B, Hs, Ws, Cs, C, act, pad, up, ps, Hd, Wd, Cd;
    int32_t pad_mode, pad2_;
} b200sp_in_apply_desc;
int b200sp_in_apply(const b200sp_in_apply_desc *d, void *stream);
/* last layer (ghiasi.py:135): out_nchw[b,c,h,w] = sigmoid(raw[b,h,w,c]*scale[b,c] + shift[b,c]), c < C */
int b200sp_in_apply_final(const float *raw, const float *scale, const float *shift, float *out_nchw,
                          int B, int H, int W, int Cs, int C, void *stream);
/* e = alpha*(noise A^T + mean) + (1-alpha)*base  (styleAugmentor.py:44-64); all fp32, A [D][D] row-major */
int b200sp_style_embed(const float *noise, const float *A, const float *mean, const float *base, float alpha,
                       float *out, int B, int D, void *stream);
/* out[b][t] = sum_d emb[b][d] W[t][d] + bias[t]: the 26 Linear(100->C) of ghiasi.py:50-51,87-90 concatenated */
int b200sp_style_linear(const float *emb, const float *W, const float *bias, float *out, int B, int D, int T, void *stream);

/* ---- SPN / AlexNet path (src/nets/spn.py:37-143, src/core/trainer.py:114-199), NHWC fp32 ------------------- */
/* Strided GEMMs (tensor-core path only): a grouped convolution (spn.py:65,73,76) is one GEMM per group over a
 * column slice of the activation.  fwd: Y[M,N] (row stride ldy) = act(X[M,K] (row stride ldx) W[N,K]^T + bias);
 * dgrad: G[M,K] = dY[M,N] (row stride lddy) W[N,K] (+skip) then the activation mask of bn->y (scale/shift/s1 NULL);
 * wgrad: dW[N,K] += dY[M,N]^T X[M,K]. */
int b200sp_gemm_fwd(const b200sp_vtensor *x, int ldx, const float *w, const float *bias, int out_act, void *y, int ldy,
                    int M, int N, int K, int dtype, void *stream);
int b200sp_gemm_dgrad(const b200sp_vtensor *dy, int lddy, const float *w, const void *skip, float scale_out, void *g,
                      const b200sp_bnbwd *bn, int M, int N, int K, int dtype, void *stream);
int b200sp_gemm_wgrad(const b200sp_vtensor *dy, int lddy, const b200sp_vtensor *x, int ldx, float *dw,
                      int M, int N, int K, int dtype, void *stream);
int b200sp_colsum_f32(const b200sp_vtensor *dy, float *out /* += */, int M, int N, int dtype, void *stream);
/* split-K FC layers for M <= 128 rows (spn.py:80-99): y_acc[M,N] += X[M,K] W[N,K]^T ; dx_acc[M,K] += dY[M,N] W[N,K]
 * (fp32 red.add: y_acc must be zero, dx_acc zero or the gradient to accumulate onto); b200sp_bias_act finishes the forward */
int b200sp_fc_fwd_splitk(const float *x, const float *w, float *y_acc, int M, int N, int K, void *stream);
int b200sp_fc_dgrad_splitk(const float *dy, const float *w, float *dx_acc, int M, int N, int K, void *stream);
int b200sp_bias_act(float *y, const float *bias, int M, int N, int relu, void *stream);
/* patch matrix of a k x k / stride / zero-pad convolution over channels [c_off, c_off+Cg) of x ([B,H,W,C] NHWC, or the
 * loader's NCHW image when nchw != 0): col[B*Ho*Wo][Kp], column (kh*k+kw)*Cg + c, zero-padded to Kp columns */
int b200sp_im2col(const float *x, float *col, int B, int H, int W, int C, int c_off, int Cg, int k, int stride, int pad,
                  int Kp, int nchw, void *stream);
/* adjoint of im2col (gather form), times the ReLU mask of act_mask (same layout as dx) when given */
int b200sp_col2im(const float *dcol, float *dx, const float *act_mask, int B, int H, int W, int C, int c_off, int Cg, int k,
                  int stride, int pad, int Kp, void *stream);
/* MaxPool2d(3,2) [+ LocalResponseNorm(size 2, alpha, beta, k=1)] (spn.py:61-62,66-67,77); pooled (may be NULL) keeps the
 * pre-LRN values and amax (may be NULL in eval) the first-maximum index 0..8 of every window for the backward */
int b200sp_pool_lrn_fwd(const float *x, float *pooled, float *out, uint8_t *amax, int B, int H, int W, int C, int lrn, float alpha,
                        float beta, void *stream);
int b200sp_pool_lrn_bwd(const float *g_out, const float *pooled, const float *x, const uint8_t *amax, float *scratch, float *dx,
                        int B, int H, int W, int C, int lrn, float alpha, float beta, int relu_mask, void *stream);
/* Dropout(p) (spn.py:81,85,92,96): counter-based mask from `seed` (not torch's RNG stream); mask saved for backward */
int b200sp_dropout_fwd(const float *x, float *out, uint8_t *mask, int64_t n, float p, uint64_t seed, void *stream);
int b200sp_dropout_bwd(float *g, const uint8_t *mask, int64_t n, float p, void *stream);
/* same mask generator with a DEVICE-resident step counter mixed into the seed (int64 *counter, advanced with b200sp_add_i64
 * once per training forward): the step can be captured in a CUDA graph and still draws a new mask per replay, every
 * forward of a reference-style loop advances it, and (seed = f(cfg.seed, rank), counter = steps taken) makes masks differ
 * across ranks and continue after a resume (spn.py:81,85,92,96 use torch's global RNG for the same purpose) */
int b200sp_dropout_fwd_ctr(const float *x, float *out, uint8_t *mask, int64_t n, float p, uint64_t seed, const int64_t *counter,
                           void *stream);
/* softmax_cross_entropy_with_logits (spn.py:37-48): loss_rows[b] = -sum_j t_bj log_softmax(z_b)_j;
 * dlogits (may be NULL) = weight/B * (softmax * sum_j t - t)  (weight: 1 for the class branch, 10 for the regress branch) */
int b200sp_soft_ce(const float *logits, const float *target, float *loss_rows, float *dlogits, int B, int N, float weight,
                   void *stream);
int b200sp_soft_ce_mean(const float *rows_c, const float *rows_r, float *loss2, int B, void *stream);
int b200sp_relu_mask(float *g, const float *a, int64_t n, void *stream);

/* ---- DANN domain classifier tail + loss (revgrad.py:75-80 AvgPool2d(7) -> Conv2d(1280,1,1); dann.py:85-92
 * binary_cross_entropy_with_logits(mean) against a constant label).  The first conv (+bias+ReLU) is
 * b200sp_pw_fwd.  h is [B,HW,C] NHWC. */
int b200sp_dann_head_fwd(const void *h, const float *w3 /*[C]*/, const float *b3 /*[1]*/, float *pooled /*[B,C]*/,
                         float *z /*[B]*/, int B, int HW, int C, int dtype, void *stream);
/* loss[0] = mean BCE-with-logits(z, label); dz[b] = (sigmoid(z_b) - label)/B * loss_scale[0] (NULL => 1) */
int b200sp_bce_logits(const float *z, float label, float *loss, float *dz, const float *loss_scale, int B, void *stream);
/* h_inout <- dL/dh = dz[b]*w3[c]/HW * (h>0)  (ReLU + AvgPool backward, in place); dw3 += , db3 += */
int b200sp_dann_head_bwd(void *h_inout, const float *dz, const float *pooled, const float *w3, float *dw3, float *db3,
                         int B, int HW, int C, int dtype, void *stream);
/* x *= s[0]*mul with s in DEVICE memory: the gradient-reversal factor -lambda (revgrad.py:52-56) */
int b200sp_scale_dev(void *x, int64_t n, const float *s, float mul, int dtype, void *stream);

/* ---- evaluation tail on the device (src/core/inference.py) ----
 * b200sp_topk_softmax: inference.py:180-181 `topW, topC = torch.topk(weights, num_neighbors, dim=1); topW = softmax(topW, 1)`.
 *   logits [B,N] fp32 -> top_w [B,k] (softmax over the k winners, descending), top_idx [B,k] int64 (ties: lowest index),
 *   top_raw [B,k] (optional, may be NULL: the un-normalised winners).  1 <= k <= 32, k <= N, N*4 <= 200 KB.
 * b200sp_kpt_denorm: inference.py:236-243 `corners2D[:,0] = x*(xmax-xmin)+xmin; [:,1] = y*(ymax-ymin)+ymin` with numpy's
 *   fp32 rounding (no FMA).  logits [B,2K] interleaved (x0,y0,x1,y1,..: park2019.py:163-164), bbox [B,4] =
 *   (xmin,xmax,ymin,ymax) in pixels -> out [B,K,2] pixels. */
int b200sp_topk_softmax(const float *logits, float *top_w, float *top_raw, int64_t *top_idx, int B, int N, int k, void *stream);
int b200sp_kpt_denorm(const float *logits, const float *bbox, float *out, int B, int K, void *stream);

/* ---- input pipeline on the device (src/datasets/transforms.py:38-246, Park2019KRNDataset.py:86-87) ----
 * One batch of raw 8-bit frames [B,H,W,C] (C = 1 grey, replicated to RGB like `.convert('RGB')`, or 3) already in HBM ->
 * fp32 NCHW [B,3,oh,ow] in [0,1]: crop to (x0,x1,y0,y1), Pillow BILINEAR resize (8-bit two-pass resampler, bit-exact),
 * ToTensor, quarter-turn rotation, flip, clamp(a*x+b), clamp(x + N(0, noise_std)).  The random DECISIONS are made by the
 * host (datasets/transforms.py follows the reference's distributions) and arrive as one b200sp_aug per image. */
typedef struct b200sp_aug {
    int32_t x0, x1, y0, y1;   /* crop box in frame pixels, [x0,x1) x [y0,y1)  (RandomCrop :147-152 / ResizeCrop :181-184) */
    int32_t rot;              /* 0..3 quarter turns counter-clockwise (Rotate :40-44: angle = 90*rot) */
    int32_t flip;             /* 0 none, 1 horizontal, 2 vertical (Flip :61-70) */
    int32_t bc;               /* 1: apply clamp(a*image + b, 0, 1) (BrightnessContrast :97) */
    float a, b;
    float noise_std;          /* > 0: clamp(image + noise_std*N(0,1), 0, 1) (GaussianNoise :110-111), device RNG keyed by seed */
    uint32_t seed;
    int32_t pad_;
} b200sp_aug;
/* aug: DEVICE array [B].  coef: int32 scratch [B*2*max(oh,ow)*(2+ks)].  tmp: uint8 scratch [B*tmp_rows*ow*C].
 * ks >= 2*ceil(max(crop extent / output extent, 1)) + 1 taps and tmp_rows >= max crop height are the caller's sizing; a box
 * that violates them (or is empty) sets bit 0/1 of *status (DEVICE int, may be NULL) and yields zeros / truncated taps
 * instead of writing out of bounds.  any_odd_rot: some rot is 1 or 3 (then oh must equal ow). */
int b200sp_input_pipeline(const uint8_t *frames, int B, int H, int W, int C, const b200sp_aug *aug, int32_t *coef, uint8_t *tmp,
                          int tmp_rows, int ks, float *out, int oh, int ow, int any_odd_rot, int *status, void *stream);
/* keypoints [B,2,K] in frame pixels -> the crop's [0,1] frame (normalize=1, RandomCrop :156-159) -> Rotate/Flip bookkeeping */
int b200sp_kpt_augment(const float *kpt_pix, const b200sp_aug *aug, float *out, int B, int K, int normalize, void *stream);

/* ---- optimizer (build.py:72-74 torch.optim.AdamW; trainer.py:97 clip_grad_norm_) ---- */
typedef struct b200sp_adamw_hp {   /* lives in DEVICE memory so CUDA graphs can replay */
    float lr, beta1, beta2, eps, weight_decay, max_norm, clip_value, grad_scale;
    int32_t step;                  /* incremented by b200sp_adamw_step */
    int32_t clip_mode;             /* 0 none, 1 global L2 norm (KRN/DANN), 2 value (SPN) */
    double sqnorm;                 /* scratch: sum g^2 (zero on entry/exit) */
    float last_norm;               /* out: total grad norm of the last step */
    float pad_;
} b200sp_adamw_hp;
int b200sp_grad_sqnorm(const float *g, int64_t n, b200sp_adamw_hp *hp, void *stream);
int b200sp_adamw_step(float *p, const float *g, float *m, float *v, void *p_lowp /*bf16 copy or NULL*/,
                      int64_t n, b200sp_adamw_hp *hp, void *stream);
/* The other --optimizer choices (build.py:63-71: torch.optim.SGD(momentum) / RMSprop(alpha=cfg.momentum) /
 * Adam(betas=(cfg.momentum, .999)), all with coupled L2 weight decay), as ONE launch over the flat buffer after the same
 * clip.  hp->beta1 carries momentum (SGD) / alpha (RMSprop) / beta1 (Adam); s1 = momentum_buffer / square_avg / exp_avg,
 * s2 = exp_avg_sq (Adam only, else NULL).  kind == B200SP_OPT_ADAMW forwards to b200sp_adamw_step. */
enum { B200SP_OPT_ADAMW = 0, B200SP_OPT_SGD = 1, B200SP_OPT_RMSPROP = 2, B200SP_OPT_ADAM = 3 };
int b200sp_optim_step(int kind, float *p, const float *g, float *s1, float *s2, void *p_lowp /*bf16 copy or NULL*/,
                      int64_t n, b200sp_adamw_hp *hp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SP_H */
