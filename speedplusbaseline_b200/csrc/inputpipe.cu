// inputpipe.cu -- the per-sample input pipeline of the reference on the device.
// Replaces src/datasets/transforms.py (RandomCrop/ResizeCrop :114-191 = PIL crop + Pillow 8-bit BILINEAR resize,
// ToTensor :192-196, Rotate :38-57, Flip :59-72, BrightnessContrast :74-99, GaussianNoise :101-112) for a whole batch of
// raw 8-bit frames already in HBM: three launches instead of 48 x (PIL + 5 torch ops) in DataLoader workers.
//
//   1. resample_coeffs_kernel   Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) per image and
//                               axis, in double with explicitly un-fused IEEE operations (bit-identical to the C code):
//                               window start, tap count and 22-bit fixed-point taps per output index.
//   2. resample_h_kernel        horizontal pass over the crop rows -> uint8 temp [B, rows, ow, C]   (integer arithmetic)
//   3. resample_v_aug_kernel    vertical pass -> uint8 -> /255 (ToTensor) -> destination index of the quarter-turn
//                               rotation + flip -> a*x+b clamp -> + noise clamp -> fp32 NCHW (grey replicated to RGB).
// Integer/byte work, HBM-bound: crop read once, 12 B written per output pixel.  The resize is bit-exact against Pillow; the
// Gaussian noise comes from a counter-based generator (not torch's stream: statistical parity only, like Dropout).
#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// coefficient table of one (image, axis): [n_out][2 + ks] int32 = {first, count, taps...}
__global__ void resample_coeffs_kernel(const b200sp_aug* __restrict__ aug, int32_t* __restrict__ coef, int oh, int ow, int ks,
                                       int tmp_rows, int* __restrict__ status) {
    const int b = blockIdx.x, axis = blockIdx.y;                  // axis 0: x (ow outputs), 1: y (oh outputs)
    const b200sp_aug g = aug[b];
    const int in_size = axis == 0 ? g.x1 - g.x0 : g.y1 - g.y0;
    const int n_out = axis == 0 ? ow : oh, n_max = ow > oh ? ow : oh;
    int32_t* tab = coef + ((size_t)(b * 2 + axis) * n_max) * (2 + ks);
    if (in_size <= 0 || g.y1 - g.y0 > tmp_rows) {
        if (threadIdx.x == 0 && status) atomicOr(status, 1);
        for (int xx = threadIdx.x; xx < n_out; xx += blockDim.x) { tab[(size_t)xx * (2 + ks)] = 0; tab[(size_t)xx * (2 + ks) + 1] = 0; }
        return;
    }
    const double scale = __ddiv_rn((double)in_size, (double)n_out);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                           // BILINEAR support 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    for (int xx = threadIdx.x; xx < n_out; xx += blockDim.x) {
        int32_t* row = tab + (size_t)xx * (2 + ks);
        const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
        int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
        if (lo < 0) lo = 0;
        int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
        if (hi > in_size) hi = in_size;
        int n = hi - lo;
        if (n > ks) { n = ks; if (status) atomicOr(status, 2); }
        double ww = 0.0;
        for (int x = 0; x < n; ++x) {
            double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
            if (t < 0.0) t = -t;
            ww = __dadd_rn(ww, t < 1.0 ? __dsub_rn(1.0, t) : 0.0);
        }
        for (int x = 0; x < n; ++x) {
            double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
            if (t < 0.0) t = -t;
            double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double v = __dmul_rn(w, (double)(1 << PRECISION_BITS));
            row[2 + x] = w < 0.0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
        }
        row[0] = lo;
        row[1] = n;
    }
}

// horizontal pass: thread = (crop row, output column); grid (ceil(rows*ow / 256), B)
template <int C>
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ frames, const b200sp_aug* __restrict__ aug,
                                                         const int32_t* __restrict__ coef, uint8_t* __restrict__ tmp,
                                                         int H, int W, int oh, int ow, int ks, int tmp_rows) {
    const int b = blockIdx.y;
    const b200sp_aug g = aug[b];
    const int rows = g.y1 - g.y0;
    if (rows > tmp_rows) return;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= rows * ow) return;
    const int y = idx / ow, ox = idx - y * ow;
    const int n_max = ow > oh ? ow : oh;
    const int32_t* row = coef + (((size_t)(b * 2 + 0) * n_max) + ox) * (2 + ks);
    const int lo = row[0], n = row[1];
    const uint8_t* src = frames + (((size_t)b * H + (g.y0 + y)) * W + (g.x0 + lo)) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < n; ++x) {
        const int k = row[2 + x];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)src[x * C + c] * k;
    }
    uint8_t* dst = tmp + (((size_t)b * tmp_rows + y) * ow + ox) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int v = acc[c] >> PRECISION_BITS;
        dst[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

// counter-based N(0,1): splitmix64 of (seed, element index) -> two uniforms -> Box-Muller
__device__ __forceinline__ float gauss_noise(uint32_t seed, uint64_t idx) {
    uint64_t z = ((uint64_t)seed << 32) ^ (idx * 0x9E3779B97F4A7C15ull) ^ 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u1 = ((uint32_t)(z >> 40) + 1u) * (1.0f / 16777216.0f);        // (0, 1]
    const float u2 = (uint32_t)(z & 0xFFFFFFu) * (1.0f / 16777216.0f);         // [0, 1)
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// where source pixel (sy, sx) of an n x n (or oh x ow when rot is even) image lands after torch.rot90(img, rot, (1,2))
// followed by the flip (1: horizontal, 2: vertical).  Mirrored by datasets/transforms.py:dest_index (CPU-tested).
__device__ __forceinline__ void dest_index(int sy, int sx, int oh, int ow, int rot, int flip, int& i, int& j) {
    if (rot == 1)      { i = ow - 1 - sx; j = sy; }
    else if (rot == 2) { i = oh - 1 - sy; j = ow - 1 - sx; }
    else if (rot == 3) { i = sx;          j = oh - 1 - sy; }
    else               { i = sy;          j = sx; }
    if (flip == 1) j = ow - 1 - j;             // after an odd rotation the image is still oh x ow because oh == ow is enforced
    else if (flip == 2) i = oh - 1 - i;
}

// vertical pass + ToTensor + augmentation: thread = (output row, output column) in SOURCE orientation; grid (.., B)
template <int C>
__global__ void __launch_bounds__(256) resample_v_aug_kernel(const uint8_t* __restrict__ tmp, const b200sp_aug* __restrict__ aug,
                                                             const int32_t* __restrict__ coef, float* __restrict__ out,
                                                             int oh, int ow, int ks, int tmp_rows) {
    const int b = blockIdx.y;
    const b200sp_aug g = aug[b];
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= oh * ow) return;
    const int oy = idx / ow, ox = idx - oy * ow;
    const int n_max = ow > oh ? ow : oh;
    const int32_t* row = coef + (((size_t)(b * 2 + 1) * n_max) + oy) * (2 + ks);
    const int lo = row[0], n = row[1];
    int i, j;
    dest_index(oy, ox, oh, ow, g.rot, g.flip, i, j);
    float* dst = out + (size_t)b * 3 * oh * ow + (size_t)i * ow + j;
    if (g.y1 - g.y0 > tmp_rows || g.y1 <= g.y0 || g.x1 <= g.x0) {          // flagged by the coefficient kernel: defined output
        dst[0] = 0.f; dst[(size_t)oh * ow] = 0.f; dst[(size_t)2 * oh * ow] = 0.f;
        return;
    }
    const uint8_t* src = tmp + (((size_t)b * tmp_rows + lo) * ow + ox) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < n; ++y) {
        const int k = row[2 + y];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)src[(size_t)y * ow * C + c] * k;
    }
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
        int v = acc[C == 1 ? 0 : c3] >> PRECISION_BITS;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        float f = __fdiv_rn((float)v, 255.0f);                              // ToTensor: uint8 -> float32 .div(255)
        if (g.bc) f = fminf(fmaxf(__fadd_rn(__fmul_rn(g.a, f), g.b), 0.f), 1.f);       // clamp(a*image + b, 0, 1): two roundings like torch
        if (g.noise_std > 0.f) {
            const uint64_t e = ((uint64_t)b * 3 + c3) * (uint64_t)(oh * ow) + (uint64_t)i * ow + j;
            f = fminf(fmaxf(__fadd_rn(f, __fmul_rn(gauss_noise(g.seed, e), g.noise_std)), 0.f), 1.f);
        }
        dst[(size_t)c3 * oh * ow] = f;
    }
}

// keypoints: pixel coordinates -> crop frame [0,1] (RandomCrop :156-159) -> Rotate / Flip bookkeeping (:46-55, :63-70)
__global__ void kpt_augment_kernel(const float* __restrict__ kin, const b200sp_aug* __restrict__ aug, float* __restrict__ kout,
                                   int B, int K, int normalize) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    const int b = idx / K, q = idx - b * K;
    const b200sp_aug g = aug[b];
    float x = kin[((size_t)b * 2 + 0) * K + q], y = kin[((size_t)b * 2 + 1) * K + q];
    if (normalize) {
        x = __fdiv_rn(__fsub_rn(x, (float)g.x0), (float)(g.x1 - g.x0));
        y = __fdiv_rn(__fsub_rn(y, (float)g.y0), (float)(g.y1 - g.y0));
    }
    const float x0 = x, y0 = y;
    if (g.rot == 1)      { x = y0;                y = __fsub_rn(1.0f, x0); }
    else if (g.rot == 2) { x = __fsub_rn(1.0f, x0); y = __fsub_rn(1.0f, y0); }
    else if (g.rot == 3) { x = __fsub_rn(1.0f, y0); y = x0; }
    if (g.flip == 1) x = __fsub_rn(1.0f, x);
    else if (g.flip == 2) y = __fsub_rn(1.0f, y);
    kout[((size_t)b * 2 + 0) * K + q] = x;
    kout[((size_t)b * 2 + 1) * K + q] = y;
}

}  // namespace

extern "C" int b200sp_input_pipeline(const uint8_t* frames, int B, int H, int W, int C, const b200sp_aug* aug, int32_t* coef,
                                     uint8_t* tmp, int tmp_rows, int ks, float* out, int oh, int ow, int any_odd_rot,
                                     int* status, void* stream) {
    if (!frames || !aug || !coef || !tmp || !out || B < 0 || H < 1 || W < 1 || (C != 1 && C != 3) || oh < 1 || ow < 1 ||
        ks < 3 || tmp_rows < 1)
        return B200SP_EINVAL;
    if (any_odd_rot && oh != ow) return B200SP_EINVAL;              // quarter turns need a square output (the reference's is)
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    resample_coeffs_kernel<<<dim3(B, 2), 256, 0, st>>>(aug, coef, oh, ow, ks, tmp_rows, status);
    B200SP_COUNT_LAUNCH();
    const dim3 gh(ceil_div((long long)tmp_rows * ow, 256), B), gv(ceil_div((long long)oh * ow, 256), B);
    if (C == 1) {
        resample_h_kernel<1><<<gh, 256, 0, st>>>(frames, aug, coef, tmp, H, W, oh, ow, ks, tmp_rows);
        B200SP_COUNT_LAUNCH();
        resample_v_aug_kernel<1><<<gv, 256, 0, st>>>(tmp, aug, coef, out, oh, ow, ks, tmp_rows);
    } else {
        resample_h_kernel<3><<<gh, 256, 0, st>>>(frames, aug, coef, tmp, H, W, oh, ow, ks, tmp_rows);
        B200SP_COUNT_LAUNCH();
        resample_v_aug_kernel<3><<<gv, 256, 0, st>>>(tmp, aug, coef, out, oh, ow, ks, tmp_rows);
    }
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_kpt_augment(const float* kpt_pix, const b200sp_aug* aug, float* out, int B, int K, int normalize, void* stream) {
    if (!kpt_pix || !aug || !out || B < 0 || K < 1) return B200SP_EINVAL;
    if (B == 0) return 0;
    kpt_augment_kernel<<<ceil_div((long long)B * K, 128), 128, 0, (cudaStream_t)stream>>>(kpt_pix, aug, out, B, K, normalize);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
