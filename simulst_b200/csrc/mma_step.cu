// MMA incremental decoding step (SimulEval agents): one launch replaces the ~25 eager ops of
// MonotonicAttention.monotonic_attention_process_infer
// (codebase/modules/monotonic_multihead_attention.py:171-299).
// One CTA per (utterance, head) row; no host reads, CUDA-graph capturable.
#include "common.cuh"

namespace simulst {

constexpr int kStepThreads = 128;
constexpr int kStepWarps = kStepThreads / kWarp;

template <typename TP, typename TE>
__global__ void __launch_bounds__(kStepThreads)
mma_step_kernel(const TP* __restrict__ p_choose, const TE* __restrict__ soft_energy,
                const int32_t* __restrict__ src_lengths, int64_t* __restrict__ head_step,
                uint8_t* __restrict__ head_read, TP* __restrict__ alpha, TE* __restrict__ beta,
                int S, int mass_preservation, float fill) {
    extern __shared__ float se[];                 // soft energy row (soft attention only)
    __shared__ int ired[kStepWarps];
    __shared__ float fred[kStepWarps];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TP* p = p_choose + (size_t)r * S;
    const int len = src_lengths ? src_lengths[r] : S;
    const int max_step = mass_preservation ? len - 1 : len;
    const long long step_in = head_step[r];

    // ---- first j >= head_step with p_j >= 0.5; the stop column `max_step` always fires (:212-237)
    int first = 0x7fffffff;
    for (int j = tid; j < S; j += kStepThreads) {
        if ((long long)j >= step_in && j != max_step && to_f32<TP>(p[j]) >= 0.5f) { first = j; break; }
    }
    first = __reduce_min_sync(kFull, first);
    if (lane == 0) ired[warp] = first;
    __syncthreads();
    first = ired[0];
#pragma unroll
    for (int w = 1; w < kStepWarps; ++w) first = min(first, ired[w]);
    const int new_step = min(first, max_step);
    const int cl = max(0, min(new_step, len - 1));                       // :240-244
    const float p_i = to_f32<TP>(p[cl]);
    const bool at_stop = new_step == max_step;
    if (tid == 0) {
        head_step[r] = new_step;                                         // :253
        head_read[r] = (at_stop && p_i < 0.5f) ? 1 : 0;                  // :255-257
    }
    const bool zero_alpha = !mass_preservation && at_stop;               // :270-275
    TP* a_row = alpha + (size_t)r * S;
    for (int j = tid; j < S; j += kStepThreads)
        a_row[j] = from_f32<TP>((j == cl && !zero_alpha) ? 1.0f : 0.0f);

    if (soft_energy == nullptr) return;
    // ---- beta = softmax(soft_energy masked beyond new_step), zero when the head has not moved (:278-294)
    const TE* e_row = soft_energy + (size_t)r * S;
    TE* b_row = beta + (size_t)r * S;
    float mx = -INFINITY;
    for (int j = tid; j < S; j += kStepThreads) {
        // masked_fill writes `fill` in the energy dtype (rounds -1e8 for 16-bit types)
        const float v = (j > new_step) ? to_f32<TE>(from_f32<TE>(fill)) : to_f32<TE>(e_row[j]);
        se[j] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if (lane == 0) fred[warp] = mx;
    __syncthreads();
    mx = fred[0];
#pragma unroll
    for (int w = 1; w < kStepWarps; ++w) mx = fmaxf(mx, fred[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < S; j += kStepThreads) {
        const float ex = expf(se[j] - mx);
        se[j] = ex;
        sum += ex;
    }
    sum = warp_sum(sum);
    if (lane == 0) fred[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < kStepWarps; ++w) sum += fred[w];
    const bool zero_beta = new_step == 0;
    for (int j = tid; j < S; j += kStepThreads)
        b_row[j] = from_f32<TE>(zero_beta ? 0.f : se[j] / sum);
}

template <typename F>
static int dispatch_dtype(int dtype, F&& f) {
    switch (dtype) {
        case SIMULST_F32: return f(float{});
        case SIMULST_BF16: return f(__nv_bfloat16{});
        case SIMULST_F16: return f(__half{});
    }
    return SIMULST_E_ARG;
}

}  // namespace simulst

using namespace simulst;

extern "C" int simulst_mma_step(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                                const int32_t* src_lengths, int64_t* head_step, uint8_t* head_read,
                                void* alpha, void* beta, int R, int S, unsigned flags, void* stream) {
    if (!p_choose || !head_step || !head_read || !alpha || !valid_dtype(p_dtype)) return SIMULST_E_ARG;
    const bool soft = soft_energy != nullptr;
    if (soft && (!beta || !valid_dtype(e_dtype))) return SIMULST_E_ARG;
    if (R < 0 || S < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (R == 0 || S == 0) return SIMULST_OK;
    const int mp = (flags & SIMULST_MMA_MASS_PRESERVATION) ? 1 : 0;
    const float fill = (soft && e_dtype == SIMULST_F16) ? -1e4f : -1e8f;
    const size_t smem = soft ? (size_t)S * sizeof(float) : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(p_dtype, [&](auto tp) {
        using TP = decltype(tp);
        return dispatch_dtype(soft ? e_dtype : p_dtype, [&](auto te) {
            using TE = decltype(te);
            auto kern = mma_step_kernel<TP, TE>;
            if (smem > 48 * 1024 &&
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return (int)SIMULST_E_SHAPE;
            }
            kern<<<R, kStepThreads, smem, st>>>((const TP*)p_choose, (const TE*)soft_energy, src_lengths,
                                                head_step, head_read, (TP*)alpha, (TE*)beta, S, mp, fill);
            return check_launch();
        });
    });
}
