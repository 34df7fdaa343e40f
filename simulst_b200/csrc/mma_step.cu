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

// Vectorised variant for rows of up to 128 * 4 * NCH frames with S % 4 == 0: every thread owns NCH
// chunks of 4 consecutive frames, both rows are fetched up front in 16 / 8-byte accesses and stay in
// registers (no shared-memory staging, the energy load overlaps the first-hit search), outputs leave
// as vector stores.  Same arithmetic as the kernel above (expf, IEEE division).
template <typename TP, typename TE, int NCH>
__global__ void __launch_bounds__(kStepThreads)
mma_step_vec_kernel(const TP* __restrict__ p_choose, const TE* __restrict__ soft_energy,
                    const int32_t* __restrict__ src_lengths, int64_t* __restrict__ head_step,
                    uint8_t* __restrict__ head_read, TP* __restrict__ alpha, TE* __restrict__ beta,
                    int S, int mass_preservation, float fill) {
    __shared__ int ired[kStepWarps];
    __shared__ float fred[2][kStepWarps];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TP* p = p_choose + (size_t)r * S;
    const bool soft = soft_energy != nullptr;
    const TE* e_row = soft ? soft_energy + (size_t)r * S : nullptr;
    Pack<TP, 4> pv[NCH];
    Pack<TE, 4> ev[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j = (c * kStepThreads + tid) * 4;
        if (j < S) {
            pv[c] = *reinterpret_cast<const Pack<TP, 4>*>(p + j);
            if (soft) ev[c] = *reinterpret_cast<const Pack<TE, 4>*>(e_row + j);
        }
    }
    const int len = src_lengths ? src_lengths[r] : S;
    const int max_step = mass_preservation ? len - 1 : len;
    const long long step_in = head_step[r];
    // ---- first j >= head_step with p_j >= 0.5; the stop column `max_step` always fires (:212-237)
    int first = 0x7fffffff;
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
        const int j0 = (c * kStepThreads + tid) * 4;
        if (j0 < S) {
#pragma unroll
            for (int k = 3; k >= 0; --k) {
                const int j = j0 + k;
                if ((long long)j >= step_in && j != max_step && to_f32<TP>(pv[c].v[k]) >= 0.5f) first = j;
            }
        }
    }
    first = __reduce_min_sync(kFull, first);
    if (lane == 0) ired[warp] = first;
    __syncthreads();
    first = ired[0];
#pragma unroll
    for (int w = 1; w < kStepWarps; ++w) first = min(first, ired[w]);
    const int new_step = min(first, max_step);
    const int cl = max(0, min(new_step, len - 1));                       // :240-244
    const bool at_stop = new_step == max_step;
    if (tid == 0) {
        const float p_i = to_f32<TP>(p[cl]);
        head_step[r] = new_step;                                         // :253
        head_read[r] = (at_stop && p_i < 0.5f) ? 1 : 0;                  // :255-257
    }
    const bool zero_alpha = !mass_preservation && at_stop;               // :270-275
    TP* a_row = alpha + (size_t)r * S;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j0 = (c * kStepThreads + tid) * 4;
        if (j0 < S) {
            Pack<TP, 4> av;
#pragma unroll
            for (int k = 0; k < 4; ++k) av.v[k] = from_f32<TP>((j0 + k == cl && !zero_alpha) ? 1.0f : 0.0f);
            *reinterpret_cast<Pack<TP, 4>*>(a_row + j0) = av;
        }
    }
    if (!soft) return;
    // ---- beta = softmax(soft_energy masked beyond new_step), zero when the head has not moved (:278-294)
    const float fillv = to_f32<TE>(from_f32<TE>(fill));     // masked_fill writes `fill` in the energy dtype
    float v[NCH][4], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j0 = (c * kStepThreads + tid) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[c][k] = (j0 + k > new_step) ? fillv : to_f32<TE>(ev[c].v[k]);
            if (j0 < S) mx = fmaxf(mx, v[c][k]);
        }
    }
    mx = warp_max(mx);
    if (lane == 0) fred[0][warp] = mx;
    __syncthreads();
    mx = fred[0][0];
#pragma unroll
    for (int w = 1; w < kStepWarps; ++w) mx = fmaxf(mx, fred[0][w]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j0 = (c * kStepThreads + tid) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[c][k] = expf(v[c][k] - mx);
            if (j0 < S) sum += v[c][k];
        }
    }
    sum = warp_sum(sum);
    if (lane == 0) fred[1][warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < kStepWarps; ++w) sum += fred[1][w];
    const bool zero_beta = new_step == 0;
    TE* b_row = beta + (size_t)r * S;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int j0 = (c * kStepThreads + tid) * 4;
        if (j0 < S) {
            Pack<TE, 4> bv;
#pragma unroll
            for (int k = 0; k < 4; ++k) bv.v[k] = from_f32<TE>(zero_beta ? 0.f : v[c][k] / sum);
            *reinterpret_cast<Pack<TE, 4>*>(b_row + j0) = bv;
        }
    }
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
            // rows of up to 2048 frames with vector-legal alignment: the register-resident variant
            const size_t pa = 4 * sizeof(TP), ea = 4 * sizeof(TE);
            const bool vec = S % 4 == 0 && S <= kStepThreads * 4 * 4 &&
                             reinterpret_cast<uintptr_t>(p_choose) % pa == 0 && reinterpret_cast<uintptr_t>(alpha) % pa == 0 &&
                             (!soft || (reinterpret_cast<uintptr_t>(soft_energy) % ea == 0 &&
                                        reinterpret_cast<uintptr_t>(beta) % ea == 0));
            if (vec) {
                if (S <= kStepThreads * 4)
                    mma_step_vec_kernel<TP, TE, 1><<<R, kStepThreads, 0, st>>>(
                        (const TP*)p_choose, (const TE*)soft_energy, src_lengths, head_step, head_read, (TP*)alpha,
                        (TE*)beta, S, mp, fill);
                else if (S <= kStepThreads * 8)
                    mma_step_vec_kernel<TP, TE, 2><<<R, kStepThreads, 0, st>>>(
                        (const TP*)p_choose, (const TE*)soft_energy, src_lengths, head_step, head_read, (TP*)alpha,
                        (TE*)beta, S, mp, fill);
                else
                    mma_step_vec_kernel<TP, TE, 4><<<R, kStepThreads, 0, st>>>(
                        (const TP*)p_choose, (const TE*)soft_energy, src_lengths, head_step, head_read, (TP*)alpha,
                        (TE*)beta, S, mp, fill);
                return check_launch();
            }
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
