// Latency-loss kernels next to the expected-delay epilogue.
//
//   simulst_dal_fwd / simulst_dal_bwd   DifferentiableAverageLagging as the reference calls it at
//       codebase/criterion/mma_criterion.py:172-177 and codebase/criterion/cif_criterion.py:211-216
//       (SimulEval's latency function; restated in oracle/latency.py).
//
// The metric is a max-plus recurrence over the target axis,
//     g'(0) = g(0),  g'(i) = max(g'(i-1) + 1/gamma, g(i)),  DAL = sum_i (g'(i) - i/gamma) / |Y|,
// which SimulEval evaluates with T rounds of cat / max over the batch (3 launches per target
// step: ~400 launches and as many autograd nodes for T = 128, on tensors of a few thousand
// floats).  Here: one warp per row, the row staged in shared memory with coalesced loads, lane 0
// walks it in the reference's order (so g' is bit-identical to the sequential evaluation), the
// final sum is a warp reduction.  Backward walks the row in reverse: the gradient of g'(i)
// accumulates along the "+ 1/gamma" chain until the step where g(i) was the maximum.
#include "common.cuh"

namespace simulst {

constexpr int kDalWarps = 4;

// row staging: delays with the target padding mask applied (masked_fill(mask, 0))
__device__ __forceinline__ void dal_stage_row(const float* __restrict__ d, const uint8_t* __restrict__ m, float* row,
                                              int T, int lane) {
    for (int i = lane; i < T; i += kWarp) row[i] = (m != nullptr && m[i] != 0) ? 0.f : d[i];
    __syncwarp();
}

__global__ void __launch_bounds__(kDalWarps * kWarp)
dal_fwd_kernel(const float* __restrict__ delays, const int64_t* __restrict__ src_lens,
               const int64_t* __restrict__ ref_lens, const uint8_t* __restrict__ tmask,
               float* __restrict__ dal, int N, int T) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row = sm + (size_t)warp * T;
    for (int n = blockIdx.x * kDalWarps + warp; n < N; n += gridDim.x * kDalWarps) {
        const uint8_t* m = tmask ? tmask + (size_t)n * T : nullptr;
        dal_stage_row(delays + (size_t)n * T, m, row, T, lane);
        int pad = 0;
        if (m != nullptr)
            for (int i = lane; i < T; i += kWarp) pad += m[i] != 0;
        pad = __reduce_add_sync(kFull, pad);
        const float tgt = (float)(T - pad);
        const float src = (float)src_lens[n];
        const float gamma = (ref_lens ? (float)ref_lens[n] : tgt) / src;
        const float inv = 1.0f / gamma;
        if (lane == 0) {
            float prev = row[0];
            for (int i = 1; i < T; ++i) {
                prev = fmaxf(prev + inv, row[i]);
                row[i] = prev;
            }
        }
        __syncwarp();
        float acc = 0.f;
        for (int i = lane; i < T; i += kWarp) {
            const float v = row[i] - (float)i / gamma;
            acc += (m != nullptr && m[i] != 0) ? 0.f : v;
        }
        acc = warp_sum(acc);
        if (lane == 0) dal[n] = acc / tgt;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kDalWarps * kWarp)
dal_bwd_kernel(const float* __restrict__ delays, const int64_t* __restrict__ src_lens,
               const int64_t* __restrict__ ref_lens, const uint8_t* __restrict__ tmask,
               const float* __restrict__ g_dal, float* __restrict__ g_delays, int N, int T) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row = sm + (size_t)warp * 2 * T;      // masked delays
    float* own = row + T;                         // 1 where g(i) was the maximum of step i
    for (int n = blockIdx.x * kDalWarps + warp; n < N; n += gridDim.x * kDalWarps) {
        const uint8_t* m = tmask ? tmask + (size_t)n * T : nullptr;
        dal_stage_row(delays + (size_t)n * T, m, row, T, lane);
        int pad = 0;
        if (m != nullptr)
            for (int i = lane; i < T; i += kWarp) pad += m[i] != 0;
        pad = __reduce_add_sync(kFull, pad);
        const float tgt = (float)(T - pad);
        const float src = (float)src_lens[n];
        const float gamma = (ref_lens ? (float)ref_lens[n] : tgt) / src;
        const float inv = 1.0f / gamma;
        const float g = g_dal[n] / tgt;
        if (lane == 0) {
            float prev = row[0];
            own[0] = 1.f;
            for (int i = 1; i < T; ++i) {
                const float carried = prev + inv;
                // ties go to the carried term: torch.max over cat([carried, g(i)]) returns index 0
                const bool mine = row[i] > carried;
                own[i] = mine ? 1.f : 0.f;
                prev = mine ? row[i] : carried;
            }
            float acc = 0.f;
            for (int i = T - 1; i >= 0; --i) {
                acc += (m != nullptr && m[i] != 0) ? 0.f : g;
                const bool mine = own[i] != 0.f;
                // masked positions entered as constants (masked_fill): no gradient to the input
                row[i] = (mine && !(m != nullptr && m[i] != 0)) ? acc : 0.f;
                if (mine) acc = 0.f;
            }
        }
        __syncwarp();
        for (int i = lane; i < T; i += kWarp) g_delays[(size_t)n * T + i] = row[i];
        __syncwarp();
    }
}

}  // namespace simulst

using namespace simulst;

extern "C" {

int simulst_dal_fwd(const float* delays, const int64_t* src_lens, const int64_t* ref_lens,
                    const uint8_t* target_padding_mask, float* dal, int N, int T, void* stream) {
    if (delays == nullptr || src_lens == nullptr || dal == nullptr) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || T > SIMULST_DAL_MAX_TGT) return SIMULST_E_SHAPE;
    if (N == 0 || T == 0) return SIMULST_OK;
    const int grid = min((N + kDalWarps - 1) / kDalWarps, 148 * 8);
    const size_t smem = (size_t)kDalWarps * T * sizeof(float);
    dal_fwd_kernel<<<grid, kDalWarps * kWarp, smem, static_cast<cudaStream_t>(stream)>>>(
        delays, src_lens, ref_lens, target_padding_mask, dal, N, T);
    return check_launch();
}

int simulst_dal_bwd(const float* delays, const int64_t* src_lens, const int64_t* ref_lens,
                    const uint8_t* target_padding_mask, const float* grad_dal, float* grad_delays,
                    int N, int T, void* stream) {
    if (delays == nullptr || src_lens == nullptr || grad_dal == nullptr || grad_delays == nullptr) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || T > SIMULST_DAL_MAX_TGT) return SIMULST_E_SHAPE;
    if (N == 0 || T == 0) return SIMULST_OK;
    const int grid = min((N + kDalWarps - 1) / kDalWarps, 148 * 8);
    const size_t smem = (size_t)kDalWarps * 2 * T * sizeof(float);
    dal_bwd_kernel<<<grid, kDalWarps * kWarp, smem, static_cast<cudaStream_t>(stream)>>>(
        delays, src_lens, ref_lens, target_padding_mask, grad_dal, grad_delays, N, T);
    return check_launch();
}

}  // extern "C"
