// CTC best alignment (forced-alignment Viterbi) with the back-trace on the device -- SURVEY 8f rank 3.
//
// Replaces, in one launch,
//   codebase/criterion/best_alignment/best_alignment.cu:58-202   ctc_alignment_log_alpha_gpu_kernel
//   codebase/criterion/best_alignment/__init__.py:58-111         final-state choice + S-iteration
//                                                                Python back-trace + label translation
// used by CIFCriterion's "align" quantity loss (codebase/criterion/cif_criterion.py:240-262).
//
// The reference kernel keeps the Viterbi row in GLOBAL memory (log_alpha [B,S,2T+1] fp32, read
// back three times per state and frame), stores the arg-max predecessors as int64 [B,S,2T+1] and
// returns both to Python, which masks, arg-maxes and then walks the S frames backwards with ~6
// torch launches per frame.  Here one CTA owns a sample:
//   * the 2T+1 state row is double-buffered in shared memory, one __syncthreads per frame, the next
//     frame's emission log-probs are fetched before the barrier;
//   * predecessors are stored as one BYTE (the jump 0 / 1 / 2) in a caller-provided workspace
//     [B,S,2T+1] -- 8x less than int64, and log_alpha is never written;
//   * the final state is chosen by the reference's rule (first -inf state at the last frame, minus
//     one, modulo and clamped to the last two states; arg-max = first maximum) and the back-trace
//     runs in the same CTA: the jump table is still in L2.
// Frames t >= input_length get state 0 (the reference's arg-max over an all -inf column).
#include "common.cuh"

namespace simulst {

constexpr int kCtcMaxThreads = 1024;
constexpr int kCtcPerThread = 4;        // states per thread -> up to 4096 states (T <= 2047)

template <typename T>
__global__ void __launch_bounds__(kCtcMaxThreads)
ctc_best_alignment_kernel(const T* __restrict__ log_probs, const int64_t* __restrict__ targets, int tgt_stride,
                          const int64_t* __restrict__ input_lengths, const int64_t* __restrict__ target_lengths,
                          int blank, uint8_t* __restrict__ jumps, float* __restrict__ nll,
                          int64_t* __restrict__ states_out, int64_t* __restrict__ labels_out,
                          int N, int S, int V, int Tmax) {
    extern __shared__ float sm[];
    const int width = 2 * Tmax + 1;
    float* row0 = sm;
    float* row1 = sm + width + 2;       // two leading pad cells per row: states -1 and -2 read as -inf
    __shared__ int s_final;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int il = (int)input_lengths[b];
    const int tl = (int)target_lengths[b];
    const int n_states = 2 * tl + 1;
    const T* lp_b = log_probs + (size_t)b * V;           // frame t at + t*N*V
    const size_t frame = (size_t)N * V;
    uint8_t* jb = jumps + (size_t)b * S * width;

    // per-thread constants of its states: label, whether the +2 jump is allowed
    int lab[kCtcPerThread];
    bool three[kCtcPerThread];
#pragma unroll
    for (int k = 0; k < kCtcPerThread; ++k) {
        const int s = tid + k * nth;
        lab[k] = blank;
        three[k] = false;
        if (s < n_states && (s & 1)) {
            lab[k] = (int)targets[(size_t)b * tgt_stride + (s >> 1)];
            three[k] = s > 1 && (int)targets[(size_t)b * tgt_stride + (s >> 1) - 1] != lab[k];
        }
    }
    // t = 0 (best_alignment.cu:92-117)
    if (tid < 2) { row0[tid] = -INFINITY; row1[tid] = -INFINITY; }
    float* cur = row0 + 2;
    float* nxt = row1 + 2;
#pragma unroll
    for (int k = 0; k < kCtcPerThread; ++k) {
        const int s = tid + k * nth;
        if (s < width) {
            float v = -INFINITY;
            if (s == 0) v = to_f32<T>(lp_b[blank]);
            else if (s == 1 && tl > 0) v = to_f32<T>(lp_b[lab[k]]);
            cur[s] = v;
        }
    }
    // emissions of frame 1
    float em[kCtcPerThread];
#pragma unroll
    for (int k = 0; k < kCtcPerThread; ++k) {
        const int s = tid + k * nth;
        em[k] = (s < n_states && 1 < il) ? to_f32<T>(lp_b[frame + lab[k]]) : 0.f;
    }
    __syncthreads();
    for (int t = 1; t < il; ++t) {
        float em_next[kCtcPerThread];
#pragma unroll
        for (int k = 0; k < kCtcPerThread; ++k) {
            const int s = tid + k * nth;
            em_next[k] = (s < n_states && t + 1 < il) ? to_f32<T>(lp_b[(size_t)(t + 1) * frame + lab[k]]) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kCtcPerThread; ++k) {
            const int s = tid + k * nth;
            if (s < n_states) {
                float best = cur[s];
                int jump = 0;
                const float a1 = cur[s - 1];                    // s = 0 reads the -inf pad
                if (a1 > best) { best = a1; jump = 1; }
                if (three[k]) {
                    const float a2 = cur[s - 2];
                    if (a2 > best) { best = a2; jump = 2; }
                }
                nxt[s] = best + em[k];
                jb[(size_t)t * width + s] = (uint8_t)jump;
            } else if (s < width) {
                nxt[s] = -INFINITY;
            }
            em[k] = em_next[k];
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    // cur = Viterbi row of frame input_length - 1
    if (tid == 0) {
        // negative log-likelihood over the two final states (best_alignment.cu:187-201)
        const float l1 = cur[2 * tl];
        const float l2 = tl > 0 ? cur[2 * tl - 1] : -INFINITY;
        float m = fmaxf(l1, l2);
        if (m == -INFINITY) m = 0.f;
        if (nll) nll[b] = -(logf(expf(l1 - m) + expf(l2 - m)) + m);
        // final state (best_alignment/__init__.py:66-91)
        int first_neg = 0;
        for (int s = 0; s < width; ++s)
            if (cur[s] == -INFINITY) { first_neg = s; break; }
        int last = (first_neg - 1) % n_states;
        if (last < 0) last += n_states;                         // python remainder
        last = min(last, n_states - 2);
        // arg-max over the states [last, n_states) = FIRST maximum; all -inf: 0 (torch.argmax)
        int arg = 0;
        float best = -INFINITY;
        for (int s = max(last, 0); s < n_states; ++s) {
            const float v = cur[s];
            if (v > best) { best = v; arg = s; }
        }
        s_final = arg;
    }
    __syncthreads();
    // frames beyond the input: state 0; then the back-trace (one thread: a chain of dependent loads
    // through the jump table this CTA just wrote)
    for (int t = il + tid; t < S; t += nth) {
        states_out[(size_t)b * S + t] = 0;
        if (labels_out) labels_out[(size_t)b * S + t] = blank;
    }
    if (tid == 0) {
        __threadfence_block();
        int s = s_final;
        for (int t = il - 1; t >= 0; --t) {
            states_out[(size_t)b * S + t] = s;
            if (labels_out)
                labels_out[(size_t)b * S + t] = (s & 1) ? targets[(size_t)b * tgt_stride + (s >> 1)] : (int64_t)blank;
            if (t > 0) s -= (int)jb[(size_t)t * width + s];
        }
    }
}

}  // namespace simulst

using namespace simulst;

extern "C" {

long long simulst_ctc_workspace_bytes(int N, int S, int Tmax) {
    if (N < 0 || S < 0 || Tmax < 0) return -1;
    return (long long)N * S * (2LL * Tmax + 1);
}

int simulst_ctc_best_alignment(const void* log_probs, int dtype, const int64_t* targets, int target_stride,
                               const int64_t* input_lengths, const int64_t* target_lengths, int blank,
                               uint8_t* workspace, float* nll, int64_t* states, int64_t* labels,
                               int N, int S, int V, int Tmax, void* stream) {
    if (!log_probs || !input_lengths || !target_lengths || !workspace || !states || !valid_dtype(dtype))
        return SIMULST_E_ARG;
    if (Tmax > 0 && !targets) return SIMULST_E_ARG;
    if (blank < 0 || blank >= V) return SIMULST_E_ARG;
    if (N < 0 || S < 1 || V < 1 || Tmax < 0 || 2 * Tmax + 1 > kCtcMaxThreads * kCtcPerThread) return SIMULST_E_SHAPE;
    if (N == 0) return SIMULST_OK;
    const int width = 2 * Tmax + 1;
    int threads = 32;
    while (threads < width && threads < kCtcMaxThreads) threads <<= 1;
    const size_t smem = (size_t)2 * (width + 2) * sizeof(float);
    auto launch = [&](auto t) {
        using T = decltype(t);
        auto kern = ctc_best_alignment_kernel<T>;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<N, threads, smem, (cudaStream_t)stream>>>((const T*)log_probs, targets, target_stride, input_lengths,
                                                         target_lengths, blank, workspace, nll, states, labels,
                                                         N, S, V, Tmax);
        return check_launch();
    };
    switch (dtype) {
        case SIMULST_F32: return launch(float{});
        case SIMULST_BF16: return launch(__nv_bfloat16{});
        default: return launch(__half{});
    }
}

}  // extern "C"
