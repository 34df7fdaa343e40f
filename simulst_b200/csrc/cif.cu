// Continuous integrate-and-fire (CIF) for sm_100a: plan (scan) -> weighted segment GATHER
// (forward) -> per-frame gather + per-row suffix scan (backward).  Deterministic, no atomics on
// the data path.  Replaces cif_function (codebase/models/torch_cif/cif.py:23-196); math in
// SURVEY Appendix A.6.
//
// Per batch row, with a_s the (scaled, masked) weights and cs_s their inclusive cumsum:
//   right_s = min(floor(cs_s / beta), T)   left_s = right_{s-1} (0 for s = 0)
//   no fire (left == right):  slot left gets weight a_s
//   fire:  slot right gets rw = cs_s - right*beta, slot left gets a_s - rw - (right-left-1)*beta,
//          every slot strictly in between gets beta
// left/right are non-decreasing in s, so output slot t draws from the contiguous frame range
// [first(t), first(t+1)] with first(t) = min{s : right_s >= t}: one warp per slot streams it.
#include "common.cuh"

namespace simulst {

constexpr int kPlanThreads = 256;

// ---- firing index of frame s (shared by every kernel so that all agree bit for bit)
__device__ __forceinline__ int fire_index(float cs, float beta, int T) {
    const float q = floorf(__fdiv_rn(cs, beta));
    return (q >= (float)T) ? T : (int)q;
}

// ---------------------------------------------------------------------------- plan
template <typename TA>
__global__ void __launch_bounds__(kPlanThreads)
cif_plan_kernel(const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                const float* __restrict__ desired_sum, const int64_t* __restrict__ target_lengths,
                float* __restrict__ csum, float* __restrict__ scale_out, float* __restrict__ alpha_sum,
                int64_t* __restrict__ lengths, int* __restrict__ t_max, int S, float beta,
                unsigned* status) {
    extern __shared__ float sm[];
    float* a = sm;
    float* scratch = sm + ((S + 1) / 2) * 2;       // 8-byte aligned: holds doubles
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TA* a_row = alpha + (size_t)b * S;
    const uint8_t* m_row = mask ? mask + (size_t)b * S : nullptr;
    unsigned bits = 0u;
    // The row has only S elements, so everything here accumulates in fp64 and rounds each
    // output to fp32 once -- the rounding behaviour of torch's CPU cumsum (double accumulator),
    // which keeps the firing indices floor(csum / beta) bit-identical to the reference's except
    // at exact ties.
    double part = 0.0;
    const int per = (S + kPlanThreads - 1) / kPlanThreads;
    const int lo = min(tid * per, S), hi = min(lo + per, S);
    for (int j = lo; j < hi; ++j) {
        const float v = to_f32<TA>(a_row[j]);
        bits |= prob_bits(v);
        const float w = (m_row && m_row[j]) ? 0.f : v;
        a[j] = w;
        part += (double)w;
    }
    double* dscratch = reinterpret_cast<double*>(scratch);
    // row sum (unscaled)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(kFull, part, d);
    if (lane == 0) dscratch[warp] = part;
    __syncthreads();
    double dtot = 0.0;
#pragma unroll
    for (int w = 0; w < kPlanThreads / kWarp; ++w) dtot += dscratch[w];
    __syncthreads();
    const float tot = (float)dtot;
    float scale = 1.0f;
    if (desired_sum != nullptr) {
        scale = __fdiv_rn(desired_sum[b], tot);                 // cif.py:70
        for (int j = lo; j < hi; ++j) a[j] *= scale;            // fp32 product, as alpha * (...)
    }
    // inclusive scan (thread chunk -> warp -> block), fp64 running sums
    double run = 0.0;
    for (int j = lo; j < hi; ++j) run += (double)a[j];
    double inc = run;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        const double o = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += o;
    }
    double exc = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) exc = 0.0;
    if (lane == kWarp - 1) dscratch[warp] = inc;
    __syncthreads();
    double off = exc;
#pragma unroll
    for (int w = 0; w < kPlanThreads / kWarp; ++w)
        if (w < warp) off += dscratch[w];
    float* c_row = csum + (size_t)b * S;
    float last_c = 0.f;
    for (int j = lo; j < hi; ++j) {
        off += (double)a[j];
        last_c = (float)off;
        c_row[j] = last_c;
    }
    if (tid == 0) {
        scale_out[b] = scale;
        alpha_sum[b] = tot;
        if (target_lengths != nullptr) lengths[b] = target_lengths[b];
    }
    if (target_lengths == nullptr && hi == S && lo < S) {
        // cif.py:75, evaluated on the scan's own last element so that the row length and the
        // firing indices can never disagree
        const long long len = (long long)floorf(__fdiv_rn(last_c, beta));
        lengths[b] = len;
        atomicMax(t_max, (int)len);
    }
    flag_status(status, bits);
}

// ---------------------------------------------------------------------------- k-ary search
// smallest s in [0, S) with cs[s] / beta >= t (cs non-decreasing); S if none.  Warp-cooperative.
__device__ __forceinline__ int first_frame(const float* __restrict__ cs, int S, float beta, float t, int lane) {
    int lo = 0, hi = S;
    while (hi - lo > kWarp) {
        const int step = (hi - lo + kWarp - 1) / kWarp;
        const int pos = min(lo + (lane + 1) * step - 1, hi - 1);
        const bool ok = __fdiv_rn(cs[pos], beta) >= t;
        const unsigned m = __ballot_sync(kFull, ok);
        if (m == 0u) return hi == S ? S : hi;       // cannot happen for hi < S (invariant), kept for safety
        const int f = __ffs(m) - 1;
        const int nlo = lo + f * step;
        hi = min(lo + (f + 1) * step, hi);
        lo = nlo;
    }
    const int pos = lo + lane;
    const bool ok = pos < hi && __fdiv_rn(cs[pos], beta) >= t;
    const unsigned m = __ballot_sync(kFull, ok);
    return m ? lo + __ffs(m) - 1 : hi;
}

// weight of frame s for output slot t (l = left, r = right, both already clipped to T)
__device__ __forceinline__ float slot_weight(int t, int l, int r, float a, float cs, float beta) {
    if (l == r) return a;
    const float rw = cs - (float)r * beta;
    if (t == r) return rw;
    if (t == l) return a - rw - (float)(r - l - 1) * beta;
    return beta;
}

// ---------------------------------------------------------------------------- forward
constexpr int kFwdWarps = 4;

template <typename TX, typename TA>
__global__ void __launch_bounds__(kFwdWarps * kWarp)
cif_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
               const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
               TX* __restrict__ out, TX* __restrict__ delays, float* __restrict__ tail_weights,
               const int64_t* __restrict__ lengths, int64_t* __restrict__ lengths_out,
               int* __restrict__ t_max2,
               int B, int S, int C, int T, int T_alloc, float beta, float tail_thres, int training) {
    const int lane = threadIdx.x & 31;
    const long long slot = (long long)blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (slot >= (long long)B * T_alloc) return;
    const int b = (int)(slot / T_alloc), t = (int)(slot % T_alloc);
    const float* cs = csum + (size_t)b * S;
    const TA* a_row = alpha + (size_t)b * S;
    const uint8_t* m_row = mask ? mask + (size_t)b * S : nullptr;
    const float sc = scale[b];
    const TX* xb = x + (size_t)b * S * C;
    TX* o_row = out + ((size_t)b * T_alloc + t) * C;

    const int s_lo = first_frame(cs, S, beta, (float)t, lane);
    int s_hi = (t >= T) ? S - 1 : min(first_frame(cs, S, beta, (float)(t + 1), lane), S - 1);

    constexpr int V = 4;
    const bool vec = (C % V == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % (V * sizeof(TX)) == 0);
    float wsum = 0.f, dsum = 0.f;
    // accumulate up to 4 packs of V channels per lane per pass over the frame range
    for (int c0 = 0; c0 < C; c0 += kWarp * V * 4) {
        float acc[4][V];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < V; ++k) acc[q][k] = 0.f;
        float ws = 0.f, ds = 0.f;
        int l = (s_lo <= 0) ? 0 : fire_index(cs[s_lo - 1], beta, T);
        for (int s = s_lo; s <= s_hi; ++s) {
            const float c_s = cs[s];
            const int r = fire_index(c_s, beta, T);
            const float a = (m_row && m_row[s]) ? 0.f : to_f32<TA>(a_row[s]) * sc;
            if (t >= l && t <= r) {
                const float w = slot_weight(t, l, r, a, c_s, beta);
                ws += w;
                ds += (l != r && t != l && t != r) ? (float)(s + 1) : __fdiv_rn(w * (float)(s + 1), beta);
                const TX* xs = xb + (size_t)s * C;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = c0 + (q * kWarp + lane) * V;
                    if (c < C) {
                        float xv[V];
                        load_vec<TX, V>(xs + c, C - c, vec, 0.f, xv);
#pragma unroll
                        for (int k = 0; k < V; ++k) acc[q][k] += w * xv[k];
                    }
                }
            }
            l = r;
        }
        if (c0 == 0) { wsum = ws; dsum = ds; }
        // tail handling (inference): the slot at the row's own length holds the partial segment
        float mul = 1.0f;
        bool zero = false;
        if (!training) {
            const long long len0 = lengths[b];
            if (t == len0) {
                if (wsum >= tail_thres) mul = __fdiv_rn(beta, wsum); else zero = true;
            } else if (t > len0) {
                zero = true;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + (q * kWarp + lane) * V;
            if (c < C) {
                float ov[V];
#pragma unroll
                for (int k = 0; k < V; ++k) ov[k] = zero ? 0.f : acc[q][k] * mul;
                store_vec<TX, V>(o_row + c, C - c, vec, ov);
            }
        }
    }
    if (C == 0) return;
    if (lane == 0) {
        delays[(size_t)b * T_alloc + t] = from_f32<TX>(dsum);
        if (!training && t == lengths[b]) {
            // single writer per row: this warp owns slot len0
            tail_weights[b] = wsum;
            const long long len1 = lengths[b] + (wsum >= tail_thres ? 1 : 0);
            lengths_out[b] = len1;
            atomicMax(t_max2, (int)len1);
        }
    }
}

// ---------------------------------------------------------------------------- backward, per frame
template <typename TX, typename TA>
__global__ void __launch_bounds__(kFwdWarps * kWarp)
cif_bwd_frame_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
                     const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                     const TX* __restrict__ g_out, const TX* __restrict__ g_delay,
                     const float* __restrict__ tail_weights, const int64_t* __restrict__ len0,
                     const int64_t* __restrict__ len1,
                     TX* __restrict__ g_x, float* __restrict__ ws_gl, float* __restrict__ ws_gd,
                     int B, int S, int C, int T, int T_out, float beta, float tail_thres, int training) {
    const int lane = threadIdx.x & 31;
    const long long fr = (long long)blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (fr >= (long long)B * S) return;
    const int b = (int)(fr / S), s = (int)(fr % S);
    const float* cs = csum + (size_t)b * S;
    const float c_s = cs[s];
    const int r = fire_index(c_s, beta, T);
    const int l = (s == 0) ? 0 : fire_index(cs[s - 1], beta, T);
    const bool pad = mask && mask[(size_t)b * S + s];
    const float a = pad ? 0.f : to_f32<TA>(alpha[(size_t)b * S + s]) * scale[b];
    const TX* xs = x + ((size_t)b * S + s) * C;
    TX* gx = g_x + ((size_t)b * S + s) * C;
    const float pos = (float)(s + 1);

    // output rows that exist and were not zeroed; the tail row carries the detached upscale
    long long l0 = 0, l1 = 0;
    float tail_mul = 1.0f;
    if (!training) {
        l0 = len0[b]; l1 = len1[b];
        if (l1 > l0) tail_mul = __fdiv_rn(beta, tail_weights[b]);
    }
    constexpr int V = 4;
    const bool vec = (C % V == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_out) |
                       reinterpret_cast<uintptr_t>(g_x)) % (V * sizeof(TX)) == 0);
    float dot_l = 0.f, dot_r = 0.f;
    for (int c0 = 0; c0 < C; c0 += kWarp * V) {
        const int c = c0 + lane * V;
        float xv[V], acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { xv[k] = 0.f; acc[k] = 0.f; }
        if (c < C) load_vec<TX, V>(xs + c, C - c, vec, 0.f, xv);
        for (int t = l; t <= r; ++t) {
            if (t >= T_out) break;                          // sliced-off dump slot(s)
            float f = 1.0f;
            if (!training) {
                if (t >= l1) continue;                      // zeroed tail rows: no gradient
                if (t == l0 && l1 > l0) f = tail_mul;
            }
            const float w = slot_weight(t, l, r, a, c_s, beta);
            if (c < C) {
                float gv[V];
                load_vec<TX, V>(g_out + ((size_t)b * T_out + t) * C + c, C - c, vec, 0.f, gv);
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    acc[k] += (w * f) * gv[k];
                    d += gv[k] * xv[k];
                }
                if (t == l) dot_l += d * f;
                if (t == r) dot_r += d * f;
            }
        }
        if (c < C) store_vec<TX, V>(gx + c, C - c, vec, acc);
    }
    dot_l = warp_sum(dot_l);
    dot_r = warp_sum(dot_r);
    if (lane == 0) {
        // delay is sliced but never zeroed nor rescaled (cif.py:183-188)
        const float gd_l = (g_delay && l < T_out) ? to_f32<TX>(g_delay[(size_t)b * T_out + l]) : 0.f;
        const float gd_r = (g_delay && r < T_out) ? to_f32<TX>(g_delay[(size_t)b * T_out + r]) : 0.f;
        const float g_lw = dot_l + __fdiv_rn(gd_l * pos, beta);
        const float g_rw = dot_r + __fdiv_rn(gd_r * pos, beta);
        ws_gl[(size_t)b * S + s] = g_lw;
        ws_gd[(size_t)b * S + s] = (r > l) ? (g_rw - g_lw) : 0.f;
    }
}

// ---------------------------------------------------------------------------- backward, per row
template <typename TA>
__global__ void __launch_bounds__(kPlanThreads)
cif_bwd_alpha_kernel(const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                     const float* __restrict__ scale, const float* __restrict__ alpha_sum,
                     const float* __restrict__ g_alpha_sum, const float* __restrict__ ws_gl,
                     const float* __restrict__ ws_gd, TA* __restrict__ g_alpha, int S, int training) {
    extern __shared__ float sm[];
    float* ga = sm;
    float* scratch = sm + S;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (S + kPlanThreads - 1) / kPlanThreads;
    const int lo = min(tid * per, S), hi = min(lo + per, S);
    // inclusive suffix scan of gd
    float run = 0.f;
    for (int j = hi - 1; j >= lo; --j) { run += ws_gd[(size_t)b * S + j]; ga[j] = run; }
    const float inc = warp_incl_suffix(run, lane);
    const float exc = lane_next(inc, lane, 0.f);
    if (lane == 0) scratch[warp] = inc;
    __syncthreads();
    float off = 0.f;
#pragma unroll
    for (int w = 0; w < kPlanThreads / kWarp; ++w)
        if (w > warp) off += scratch[w];
    off += exc;
    __syncthreads();
    float dotp = 0.f;
    const uint8_t* m_row = mask ? mask + (size_t)b * S : nullptr;
    for (int j = lo; j < hi; ++j) {
        const float g = ws_gl[(size_t)b * S + j] + (off + ga[j]);
        ga[j] = g;
        const float al = (m_row && m_row[j]) ? 0.f : to_f32<TA>(alpha[(size_t)b * S + j]);
        dotp += g * al;
    }
    float extra = g_alpha_sum ? g_alpha_sum[b] : 0.f;
    float k = 1.0f, corr = 0.f;
    if (training) {
        dotp = warp_sum(dotp);
        if (lane == 0) scratch[warp] = dotp;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kPlanThreads / kWarp; ++w) tot += scratch[w];
        k = scale[b];
        corr = __fdiv_rn(tot, alpha_sum[b]);
    }
    for (int j = lo; j < hi; ++j) {
        const bool pad = m_row && m_row[j];
        const float g = k * (ga[j] - corr) + extra;
        g_alpha[(size_t)b * S + j] = from_f32<TA>(pad ? 0.f : g);
    }
}

template <typename F>
static int dispatch_t(int dtype, F&& f) {
    switch (dtype) {
        case SIMULST_F32: return f(float{});
        case SIMULST_BF16: return f(__nv_bfloat16{});
        case SIMULST_F16: return f(__half{});
    }
    return SIMULST_E_ARG;
}

}  // namespace simulst

using namespace simulst;

extern "C" {

int simulst_cif_plan(const void* alpha, int a_dtype, const uint8_t* padding_mask, const float* desired_sum,
                     const int64_t* target_lengths, float* csum, float* scale, float* alpha_sum,
                     int64_t* lengths, int* t_max, int B, int S, float beta, unsigned* status, void* stream) {
    if (!alpha || !csum || !scale || !alpha_sum || !lengths || !valid_dtype(a_dtype)) return SIMULST_E_ARG;
    if ((desired_sum == nullptr) != (target_lengths == nullptr)) return SIMULST_E_ARG;
    if (target_lengths == nullptr && t_max == nullptr) return SIMULST_E_ARG;
    if (!(beta > 0.f)) return SIMULST_E_ARG;
    if (B < 0 || S < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (B == 0 || S == 0) return SIMULST_OK;
    const size_t smem = ((size_t)S + 64 + 2) / 2 * 2 * sizeof(float);
    return dispatch_t(a_dtype, [&](auto ta) {
        using TA = decltype(ta);
        auto kern = cif_plan_kernel<TA>;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<B, kPlanThreads, smem, (cudaStream_t)stream>>>((const TA*)alpha, padding_mask, desired_sum,
                                                           target_lengths, csum, scale, alpha_sum, lengths,
                                                           t_max, S, beta, status);
        return check_launch();
    });
}

int simulst_cif_fwd(const void* input, int x_dtype, const float* csum, const float* scale, const void* alpha,
                    int a_dtype, const uint8_t* padding_mask, void* cif_out, void* delays,
                    float* tail_weights, const int64_t* lengths, int64_t* lengths_out, int* t_max2,
                    int B, int S, int C, int T, int T_alloc, float beta, float tail_thres, int training,
                    void* stream) {
    if (!input || !csum || !scale || !alpha || !cif_out || !delays || !lengths) return SIMULST_E_ARG;
    if (!valid_dtype(x_dtype) || !valid_dtype(a_dtype)) return SIMULST_E_ARG;
    if (!training && (!tail_weights || !t_max2 || !lengths_out)) return SIMULST_E_ARG;
    if (B < 0 || S < 0 || C < 0 || T < 0 || T_alloc < 0) return SIMULST_E_SHAPE;
    if (B == 0 || T_alloc == 0 || S == 0) return SIMULST_OK;
    const long long slots = (long long)B * T_alloc;
    const unsigned blocks = (unsigned)((slots + kFwdWarps - 1) / kFwdWarps);
    return dispatch_t(x_dtype, [&](auto tx) {
        using TX = decltype(tx);
        return dispatch_t(a_dtype, [&](auto ta) {
            using TA = decltype(ta);
            cif_fwd_kernel<TX, TA><<<blocks, kFwdWarps * kWarp, 0, (cudaStream_t)stream>>>(
                (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, (TX*)cif_out, (TX*)delays,
                tail_weights, lengths, lengths_out, t_max2, B, S, C, T, T_alloc, beta, tail_thres, training);
            return check_launch();
        });
    });
}

int simulst_cif_bwd(const void* input, int x_dtype, const float* csum, const float* scale, const void* alpha,
                    int a_dtype, const uint8_t* padding_mask, const void* grad_out, const void* grad_delays,
                    const float* tail_weights, const int64_t* lengths_before_tail,
                    const int64_t* lengths_after_tail, const float* alpha_sum, const float* grad_alpha_sum,
                    void* grad_input, void* grad_alpha, float* workspace, int B, int S, int C, int T,
                    int T_out, float beta, float tail_thres, int training, void* stream) {
    if (!input || !csum || !scale || !alpha || !grad_input || !grad_alpha || !workspace || !alpha_sum)
        return SIMULST_E_ARG;
    if (!grad_out && T_out > 0) return SIMULST_E_ARG;
    if (!valid_dtype(x_dtype) || !valid_dtype(a_dtype)) return SIMULST_E_ARG;
    if (!training && (!tail_weights || !lengths_before_tail || !lengths_after_tail)) return SIMULST_E_ARG;
    if (B < 0 || S < 0 || C < 0 || T < 0 || T_out < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (B == 0 || S == 0) return SIMULST_OK;
    float* ws_gl = workspace;
    float* ws_gd = workspace + (size_t)B * S;
    const long long frames = (long long)B * S;
    const unsigned blocks = (unsigned)((frames + kFwdWarps - 1) / kFwdWarps);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dispatch_t(x_dtype, [&](auto tx) {
        using TX = decltype(tx);
        return dispatch_t(a_dtype, [&](auto ta) {
            using TA = decltype(ta);
            cif_bwd_frame_kernel<TX, TA><<<blocks, kFwdWarps * kWarp, 0, st>>>(
                (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, (const TX*)grad_out,
                (const TX*)grad_delays, tail_weights, lengths_before_tail, lengths_after_tail,
                (TX*)grad_input, ws_gl, ws_gd, B, S, C, T, T_out, beta, tail_thres, training);
            return check_launch();
        });
    });
    if (rc != SIMULST_OK) return rc;
    const size_t smem = ((size_t)S + 64) * sizeof(float);
    return dispatch_t(a_dtype, [&](auto ta) {
        using TA = decltype(ta);
        auto kern = cif_bwd_alpha_kernel<TA>;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<B, kPlanThreads, smem, st>>>((const TA*)alpha, padding_mask, scale, alpha_sum, grad_alpha_sum,
                                            ws_gl, ws_gd, (TA*)grad_alpha, S, training);
        return check_launch();
    });
}

}  // extern "C"
