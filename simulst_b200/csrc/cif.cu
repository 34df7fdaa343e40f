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
#include <algorithm>
#include <initializer_list>
#include <type_traits>

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
                int64_t* __restrict__ lengths, int* __restrict__ t_max,
                int* __restrict__ seg_first, int seg_stride, int S, float beta,
                unsigned* status) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    float* a = sm;
    float* scratch = sm + ((S + 1) / 2) * 2;       // 8-byte aligned: holds doubles
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TA* a_row = alpha + (size_t)b * S;
    const uint8_t* m_row = mask ? mask + (size_t)b * S : nullptr;
    unsigned bits = 0u;
    int* f_row = seg_first + (size_t)b * seg_stride;
    for (int t = tid; t < seg_stride; t += kPlanThreads) f_row[t] = S;     // "no frame reaches slot t"
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
    // segment table: seg_first[t] = first frame whose (unclipped) firing index reaches t.  Firing
    // indices are re-derived from the STORED fp32 csum so every kernel sees the same integers.
    __syncthreads();
    {
        const float lim = (float)(seg_stride - 1);
        for (int j = lo; j < hi; ++j) {
            const float qf = fminf(floorf(__fdiv_rn(c_row[j], beta)), lim);
            const float pf = (j == 0) ? -1.0f : fminf(floorf(__fdiv_rn(c_row[j - 1], beta)), lim);
            const int q = (int)qf;
            for (int t = (int)pf + 1; t <= q; ++t) f_row[t] = j;
        }
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

// weight of frame s for output slot t (l = left, r = right, both already clipped to T)
__device__ __forceinline__ float slot_weight(int t, int l, int r, float a, float cs, float beta) {
    if (l == r) return a;
    const float rw = cs - (float)r * beta;
    if (t == r) return rw;
    if (t == l) return a - rw - (float)(r - l - 1) * beta;
    return beta;
}

}  // namespace simulst

#include "cif_tile.cuh"

namespace simulst {

// ---------------------------------------------------------------------------- forward (fallback)
constexpr int kFwdWarps = 4;
constexpr int kCifV = 4;            // channels per lane per pack
constexpr int kFwdUnroll = 4;       // frames whose loads are in flight together

// One warp per output slot (b, t).  The slot's frame range [seg_first[t], seg_first[t+1]] comes
// from the plan kernel's table; lanes first evaluate the per-frame weights of up to 32 frames
// in parallel, then the warp streams the frames (lanes over channels, NP packs of 4 channels per
// lane, kFwdUnroll frames of loads in flight).  Accumulation order is frame order: deterministic.
template <typename TX, typename TA, int NP>
__global__ void __launch_bounds__(kFwdWarps * kWarp)
cif_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
               const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
               const int* __restrict__ seg_first, int seg_stride,
               TX* __restrict__ out, TX* __restrict__ delays, float* __restrict__ tail_weights,
               const int64_t* __restrict__ lengths, int64_t* __restrict__ lengths_out,
               int* __restrict__ t_max2,
               int B, int S, int C, int T, int T_alloc, float beta, float tail_thres, int training) {
    constexpr int V = kCifV;
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
    const int* f_row = seg_first + (size_t)b * seg_stride;

    const int s_lo = (t < seg_stride) ? f_row[t] : S;
    const int s_hi = (t >= T || t + 1 >= seg_stride) ? S - 1 : min(f_row[t + 1], S - 1);
    const bool vec = (C % V == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % (V * sizeof(TX)) == 0);

    float wsum = 0.f, dsum = 0.f;
    for (int c0 = 0; c0 < max(C, 1); c0 += kWarp * V * NP) {       // (C == 0 still yields delays)
        float acc[NP][V];
#pragma unroll
        for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int k = 0; k < V; ++k) acc[q][k] = 0.f;
        float ws = 0.f, ds = 0.f;
        for (int base = s_lo; base <= s_hi; base += kWarp) {
            // ---- per-frame weight of this slot, one frame per lane
            const int s = base + lane;
            float w = 0.f, d = 0.f;
            if (s <= s_hi) {
                const float c_s = cs[s];
                const int r = fire_index(c_s, beta, T);
                const int l = (s == 0) ? 0 : fire_index(cs[s - 1], beta, T);
                const float a = (m_row && m_row[s]) ? 0.f : to_f32<TA>(a_row[s]) * sc;
                if (t >= l && t <= r) {
                    w = slot_weight(t, l, r, a, c_s, beta);
                    d = (l != r && t != l && t != r) ? (float)(s + 1) : __fdiv_rn(w * (float)(s + 1), beta);
                }
            }
            const int n = min(kWarp, s_hi - base + 1);
            for (int f0 = 0; f0 < n; f0 += kFwdUnroll) {
                float xv[kFwdUnroll][NP][V];
#pragma unroll
                for (int u = 0; u < kFwdUnroll; ++u) {
                    const TX* xs = xb + (size_t)(base + f0 + u) * C;
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const int c = c0 + (q * kWarp + lane) * V;
                        if (f0 + u < n && c < C) {
                            load_vec<TX, V>(xs + c, C - c, vec, 0.f, xv[u][q]);
                        } else {
#pragma unroll
                            for (int k = 0; k < V; ++k) xv[u][q][k] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kFwdUnroll; ++u) {
                    const float wu = __shfl_sync(kFull, w, (f0 + u) & 31);
                    const float du = __shfl_sync(kFull, d, (f0 + u) & 31);
                    if (f0 + u < n) {
                        ws += wu;
                        ds += du;
#pragma unroll
                        for (int q = 0; q < NP; ++q)
#pragma unroll
                            for (int k = 0; k < V; ++k) acc[q][k] += wu * xv[u][q][k];
                    }
                }
            }
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
        for (int q = 0; q < NP; ++q) {
            const int c = c0 + (q * kWarp + lane) * V;
            if (c < C) {
                float ov[V];
#pragma unroll
                for (int k = 0; k < V; ++k) ov[k] = zero ? 0.f : acc[q][k] * mul;
                store_vec<TX, V>(o_row + c, C - c, vec, ov);
            }
        }
    }
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
// One warp per source frame: grad_input[s] = sum over the slots l..r the frame feeds of
// weight * grad_out[slot], and the two weight gradients <grad_out[l], x_s>, <grad_out[r], x_s>.
// The x row and the grad_out rows of slots l and r (the only ones unless a frame fires more
// than once) are requested together before any of them is consumed.
template <typename TX, typename TA, int NP>
__global__ void __launch_bounds__(kFwdWarps * kWarp)
cif_bwd_frame_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
                     const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                     const TX* __restrict__ g_out, const TX* __restrict__ g_delay,
                     const float* __restrict__ tail_weights, const int64_t* __restrict__ len0,
                     const int64_t* __restrict__ len1,
                     TX* __restrict__ g_x, float* __restrict__ ws_gl, float* __restrict__ ws_gd,
                     int B, int S, int C, int T, int T_out, float beta, float tail_thres, int training) {
    constexpr int V = kCifV;
    const int lane = threadIdx.x & 31;
    const long long fr = (long long)blockIdx.x * kFwdWarps + (threadIdx.x >> 5);
    if (fr >= (long long)B * S) return;
    const int b = (int)(fr / S), s = (int)(fr % S);
    const float* cs = csum + (size_t)b * S;
    const TX* xs = x + ((size_t)b * S + s) * C;
    TX* gx = g_x + ((size_t)b * S + s) * C;
    const bool vec = (C % V == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_out) |
                       reinterpret_cast<uintptr_t>(g_x)) % (V * sizeof(TX)) == 0);
    const float c_s = cs[s];
    const float c_p = (s == 0) ? 0.f : cs[s - 1];
    const bool pad = mask && mask[(size_t)b * S + s];
    const float a_raw = to_f32<TA>(alpha[(size_t)b * S + s]);
    const float sc = scale[b];
    const int r = fire_index(c_s, beta, T);
    const int l = (s == 0) ? 0 : fire_index(c_p, beta, T);
    const float a = pad ? 0.f : a_raw * sc;
    const float pos = (float)(s + 1);

    // output rows that exist and were not zeroed; the tail row carries the detached upscale
    long long l0 = 0, l1 = 0;
    float tail_mul = 1.0f;
    if (!training) {
        l0 = len0[b]; l1 = len1[b];
        if (l1 > l0) tail_mul = __fdiv_rn(beta, tail_weights[b]);
    }
    // a slot contributes iff it was kept (t < T_out) and, in inference, not zeroed (t < l1)
    auto live = [&](int t) -> bool { return t < T_out && (training || t < l1); };
    auto factor = [&](int t) -> float { return (!training && t == l0 && l1 > l0) ? tail_mul : 1.0f; };

    float dot_l = 0.f, dot_r = 0.f;
    for (int c0 = 0; c0 < C; c0 += kWarp * V * NP) {
        float xv[NP][V], acc[NP][V], gl[NP][V], gr[NP][V];
        const bool use_l = live(l), use_r = r > l && live(r);
        const TX* gl_row = g_out + ((size_t)b * T_out + l) * C;
        const TX* gr_row = g_out + ((size_t)b * T_out + r) * C;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int c = c0 + (q * kWarp + lane) * V;
#pragma unroll
            for (int k = 0; k < V; ++k) { xv[q][k] = 0.f; acc[q][k] = 0.f; gl[q][k] = 0.f; gr[q][k] = 0.f; }
            if (c < C) {
                load_vec<TX, V>(xs + c, C - c, vec, 0.f, xv[q]);
                if (use_l) load_vec<TX, V>(gl_row + c, C - c, vec, 0.f, gl[q]);
                if (use_r) load_vec<TX, V>(gr_row + c, C - c, vec, 0.f, gr[q]);
            }
        }
        // slot t with its grad_out pack gv: acc += (w f) gv ; d = <gv, x>
        auto apply = [&](int t, const float (&gv)[NP][V]) {
            const float f = factor(t);
            const float w = slot_weight(t, l, r, a, c_s, beta);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (c0 + (q * kWarp + lane) * V < C) {
                    float d = 0.f;
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        acc[q][k] += (w * f) * gv[q][k];
                        d += gv[q][k] * xv[q][k];
                    }
                    if (t == l) dot_l += d * f;
                    if (t == r) dot_r += d * f;
                }
            }
        };
        if (use_l) apply(l, gl);
        for (int t = l + 1; t < r; ++t) {                   // frames that fire more than once
            if (!live(t)) { if (t >= T_out) break; continue; }
            float gm[NP][V];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c = c0 + (q * kWarp + lane) * V;
#pragma unroll
                for (int k = 0; k < V; ++k) gm[q][k] = 0.f;
                if (c < C) load_vec<TX, V>(g_out + ((size_t)b * T_out + t) * C + c, C - c, vec, 0.f, gm[q]);
            }
            apply(t, gm);
        }
        if (use_r) apply(r, gr);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int c = c0 + (q * kWarp + lane) * V;
            if (c < C) store_vec<TX, V>(gx + c, C - c, vec, acc[q]);
        }
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
    pdl_launch_dependents();
    pdl_wait();
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

// packs of 4 channels per lane needed to cover C channels in one pass (at most 4: 512 channels)
template <typename F>
static int dispatch_np(int C, F&& f) {
    if (C <= kWarp * kCifV) return f(std::integral_constant<int, 1>{});
    if (C <= 2 * kWarp * kCifV) return f(std::integral_constant<int, 2>{});
    return f(std::integral_constant<int, 4>{});
}

// The tile kernels need every row (C elements) to be a whole number of 16-byte units and the
// tensors to start on a 16-byte boundary (bulk-copy alignment), and C <= 512 (one pass).
static bool tile_ok(int C, size_t esize, std::initializer_list<const void*> ptrs) {
    if (C <= 0 || C > 4 * kWarp * 4 || ((size_t)C * esize) % 16 != 0) return false;
    for (const void* p : ptrs)
        if (p != nullptr && reinterpret_cast<uintptr_t>(p) % 16 != 0) return false;
    return true;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return SIMULST_E_LAUNCH;
    }
    return SIMULST_OK;
}

}  // namespace simulst

using namespace simulst;

// development / test switch: route simulst_cif_fwd/_bwd through the per-warp fallback kernels
static std::atomic<int> g_cif_force_fallback{0};   // tuning knobs: relaxed atomics, see mma_api.cu
// tuning overrides (0 = automatic): frames per forward tile chunk, frames per backward tile
static std::atomic<int> g_cif_fc{0}, g_cif_fr{0};

extern "C" {

int simulst_cif_set_tile_rows(int fwd_chunk_frames, int bwd_tile_frames) {
    if (fwd_chunk_frames < 0 || bwd_tile_frames < 0 || bwd_tile_frames % kTileWarps != 0) return SIMULST_E_ARG;
    g_cif_fc = fwd_chunk_frames;
    g_cif_fr = bwd_tile_frames;
    return SIMULST_OK;
}

long long simulst_cif_workspace_bytes(int B, int S) {
    if (B < 0 || S < 0) return SIMULST_E_SHAPE;
    return 2ll * B * S * (long long)sizeof(float);      // two fp32 rows per batch row (simulst_cif_bwd)
}

int simulst_cif_seg_stride(int training, int t_cap, int S, float beta) {
    if (t_cap < 0 || S < 0 || !(beta > 0.f)) return SIMULST_E_ARG;
    // slots 0..T+1 (training) / up to floor(S/beta)+1 fires plus the two sentinels (inference)
    return training ? t_cap + 2 : (int)((float)S / beta) + 3;
}

int simulst_cif_set_tile(int enable) {
    g_cif_force_fallback = enable ? 0 : 1;
    return SIMULST_OK;
}

int simulst_cif_plan(const void* alpha, int a_dtype, const uint8_t* padding_mask, const float* desired_sum,
                     const int64_t* target_lengths, float* csum, float* scale, float* alpha_sum,
                     int64_t* lengths, int* t_max, int32_t* seg_first, int seg_stride, int B, int S, float beta,
                     unsigned* status, void* stream) {
    if (!alpha || !csum || !scale || !alpha_sum || !lengths || !seg_first || !valid_dtype(a_dtype))
        return SIMULST_E_ARG;
    if (seg_stride < 2) return SIMULST_E_SHAPE;
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
        launch_pdl(kern, B, kPlanThreads, smem, (cudaStream_t)stream, (const TA*)alpha, padding_mask, desired_sum,
                   target_lengths, csum, scale, alpha_sum, lengths, t_max, seg_first, seg_stride, S, beta, status);
        return check_launch();
    });
}

int simulst_cif_fwd(const void* input, int x_dtype, const float* csum, const float* scale, const void* alpha,
                    int a_dtype, const uint8_t* padding_mask, const int32_t* seg_first, int seg_stride,
                    void* cif_out, void* delays,
                    float* tail_weights, const int64_t* lengths, int64_t* lengths_out, int* t_max2,
                    int B, int S, int C, int T, int T_alloc, float beta, float tail_thres, int training,
                    void* stream) {
    if (!input || !csum || !scale || !alpha || !cif_out || !delays || !lengths || !seg_first) return SIMULST_E_ARG;
    if (seg_stride < 2) return SIMULST_E_SHAPE;
    if (!valid_dtype(x_dtype) || !valid_dtype(a_dtype)) return SIMULST_E_ARG;
    if (!training && (!tail_weights || !t_max2 || !lengths_out)) return SIMULST_E_ARG;
    if (B < 0 || S < 0 || C < 0 || T < 0 || T_alloc < 0) return SIMULST_E_SHAPE;
    if (B == 0 || T_alloc == 0 || S == 0) return SIMULST_OK;
    const long long slots = (long long)B * T_alloc;
    const unsigned blocks = (unsigned)((slots + kFwdWarps - 1) / kFwdWarps);
    const bool tile = tile_ok(C, dtype_size(x_dtype), {input, cif_out}) && !g_cif_force_fallback;
    return dispatch_t(x_dtype, [&](auto tx) {
        using TX = decltype(tx);
        return dispatch_t(a_dtype, [&](auto ta) {
            using TA = decltype(ta);
            return dispatch_np(C, [&](auto np) {
                constexpr int NP = decltype(np)::value;
                if (tile) {
                    const size_t row_bytes = (size_t)C * sizeof(TX);
                    const int fc_forced = g_cif_fc.load(std::memory_order_relaxed);
                    const int FC = fc_forced > 0 ? fc_forced : (int)std::max<size_t>(8, 48 * 1024 / row_bytes);
                    const size_t smem = kTileHeader + (size_t)FC * row_bytes;
                    auto kern = cif_fwd_tile_kernel<TX, TA, NP>;
                    if (int rc = set_smem(kern, smem)) return rc;
                    const unsigned grid = (unsigned)((long long)B * ((T_alloc + kTileWarps - 1) / kTileWarps));
                    launch_pdl(kern, grid, kTileThreads, smem, (cudaStream_t)stream,
                        (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, seg_first, seg_stride,
                        (TX*)cif_out, (TX*)delays, tail_weights, lengths, lengths_out, t_max2, B, S, C, T,
                        T_alloc, beta, tail_thres, training, FC);
                    return check_launch();
                }
                cif_fwd_kernel<TX, TA, NP><<<blocks, kFwdWarps * kWarp, 0, (cudaStream_t)stream>>>(
                    (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, seg_first, seg_stride,
                    (TX*)cif_out, (TX*)delays, tail_weights, lengths, lengths_out, t_max2, B, S, C, T,
                    T_alloc, beta, tail_thres, training);
                return check_launch();
            });
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
    const bool tile = tile_ok(C, dtype_size(x_dtype), {input, grad_out, grad_input}) && !g_cif_force_fallback;
    int rc = dispatch_t(x_dtype, [&](auto tx) {
        using TX = decltype(tx);
        return dispatch_t(a_dtype, [&](auto ta) {
            using TA = decltype(ta);
            return dispatch_np(C, [&](auto np) {
                constexpr int NP = decltype(np)::value;
                if (tile) {
                    const size_t row_bytes = (size_t)C * sizeof(TX);
                    int FR = (int)(32 * 1024 / row_bytes) / kTileWarps * kTileWarps;
                    FR = std::max(kTileWarps, std::min(64, FR));
                    if (const int fr_forced = g_cif_fr.load(std::memory_order_relaxed); fr_forced > 0) FR = fr_forced;
                    const int GR = FR / 2 + 2;
                    const size_t smem = kTileHeader + (size_t)(FR + GR) * row_bytes;
                    auto kern = cif_bwd_tile_kernel<TX, TA, NP>;
                    if (int rc2 = set_smem(kern, smem)) return rc2;
                    const unsigned grid = (unsigned)((long long)B * ((S + FR - 1) / FR));
                    launch_pdl(kern, grid, kTileThreads, smem, st,
                        (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, (const TX*)grad_out,
                        (const TX*)grad_delays, tail_weights, lengths_before_tail, lengths_after_tail,
                        (TX*)grad_input, ws_gl, ws_gd, B, S, C, T, T_out, beta, tail_thres, training, FR, GR);
                    return check_launch();
                }
                cif_bwd_frame_kernel<TX, TA, NP><<<blocks, kFwdWarps * kWarp, 0, st>>>(
                    (const TX*)input, csum, scale, (const TA*)alpha, padding_mask, (const TX*)grad_out,
                    (const TX*)grad_delays, tail_weights, lengths_before_tail, lengths_after_tail,
                    (TX*)grad_input, ws_gl, ws_gd, B, S, C, T, T_out, beta, tail_thres, training);
                return check_launch();
            });
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
        launch_pdl(kern, B, kPlanThreads, smem, st, (const TA*)alpha, padding_mask, scale, alpha_sum, grad_alpha_sum,
                   ws_gl, ws_gd, (TA*)grad_alpha, S, training);
        return check_launch();
    });
}

}  // extern "C"
