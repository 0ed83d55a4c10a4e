// postproc.cu -- on-device tail of the evaluation loops.
// Replaces, per batch instead of per image, the torch/numpy call sites of src/core/inference.py:
//   :180-181  topWeights, topClasses = torch.topk(weights, cfg.num_neighbors, dim=1); topWeights = softmax(topWeights, 1)
//   :236-243  corners2D[:, 0] = x * (xmax - xmin) + xmin;  corners2D[:, 1] = y * (ymax - ymin) + ymin   (numpy fp32)
// so one evaluation batch needs ONE small device->host copy (B*k*12 bytes, or B*K*8 bytes) instead of the logits of every
// image.  EPnP / SPEED metrics stay on the CPU (out of scope).
#include <climits>
#include "common.cuh"

namespace {

constexpr int TK_T = 256, TK_MAXK = 32;

// one CTA per row: the row is staged in shared memory once, then k block-wide argmax passes pick the winners in
// descending order (ties: lowest index).  Selected entries are overwritten with NaN, which the scan skips, so genuine
// -inf logits stay selectable; NaN logits are never selected (torch.topk would rank them first -- not supported here).
__global__ void __launch_bounds__(TK_T) topk_softmax_kernel(const float* __restrict__ w, float* __restrict__ top_w,
                                                            float* __restrict__ top_raw, long long* __restrict__ top_idx,
                                                            int N, int k) {
    extern __shared__ float row[];
    __shared__ float s_v[TK_T / 32];
    __shared__ int s_i[TK_T / 32];
    __shared__ float s_top[TK_MAXK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = w + (size_t)blockIdx.x * N;
    for (int i = tid; i < N; i += TK_T) row[i] = __ldg(src + i);
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        float bv = -__int_as_float(0x7f800000);
        int bi = INT_MAX;
        for (int i = tid; i < N; i += TK_T) {
            const float v = row[i];
            if (v == v && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < TK_T / 32; ++q)
                if (s_v[q] > bv || (s_v[q] == bv && s_i[q] < bi)) { bv = s_v[q]; bi = s_i[q]; }
            s_top[j] = bv;
            top_idx[(size_t)blockIdx.x * k + j] = bi == INT_MAX ? -1 : bi;
            if (top_raw) top_raw[(size_t)blockIdx.x * k + j] = bv;
            if (bi != INT_MAX) row[bi] = __int_as_float(0x7fc00000);
        }
        __syncthreads();
    }
    if (tid == 0) {                      // softmax over the k winners (s_top[0] is the maximum)
        const float m = s_top[0];
        float sum = 0.f;
        for (int j = 0; j < k; ++j) sum += expf(s_top[j] - m);
        for (int j = 0; j < k; ++j) top_w[(size_t)blockIdx.x * k + j] = expf(s_top[j] - m) / sum;
    }
}

// numpy evaluates `x * (xmax - xmin) + xmin` on float32 with one rounding per operation: no FMA contraction here
__global__ void kpt_denorm_kernel(const float* __restrict__ logits, const float* __restrict__ bbox, float* __restrict__ out,
                                  int B, int K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * K) return;
    const int b = idx / K, q = idx - b * K;
    const float x = logits[(size_t)b * 2 * K + 2 * q], y = logits[(size_t)b * 2 * K + 2 * q + 1];
    const float xmin = bbox[b * 4 + 0], xmax = bbox[b * 4 + 1], ymin = bbox[b * 4 + 2], ymax = bbox[b * 4 + 3];
    out[(size_t)idx * 2 + 0] = __fadd_rn(__fmul_rn(x, __fsub_rn(xmax, xmin)), xmin);
    out[(size_t)idx * 2 + 1] = __fadd_rn(__fmul_rn(y, __fsub_rn(ymax, ymin)), ymin);
}

}  // namespace

extern "C" int b200sp_topk_softmax(const float* logits, float* top_w, float* top_raw, int64_t* top_idx, int B, int N, int k,
                                   void* stream) {
    if (!logits || !top_w || !top_idx || B < 0 || N < 1 || k < 1 || k > TK_MAXK || k > N) return B200SP_EINVAL;
    if ((size_t)N * 4 > 200 * 1024) return B200SP_ENOSYS;
    if (B == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(topk_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    topk_softmax_kernel<<<B, TK_T, (size_t)N * 4, (cudaStream_t)stream>>>(logits, top_w, top_raw, (long long*)top_idx, N, k);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}

extern "C" int b200sp_kpt_denorm(const float* logits, const float* bbox, float* out, int B, int K, void* stream) {
    if (!logits || !bbox || !out || B < 0 || K < 1) return B200SP_EINVAL;
    if (B == 0) return 0;
    kpt_denorm_kernel<<<ceil_div((long long)B * K, 128), 128, 0, (cudaStream_t)stream>>>(logits, bbox, out, B, K);
    B200SP_COUNT_LAUNCH();
    B200SP_RETURN_LAST();
}
