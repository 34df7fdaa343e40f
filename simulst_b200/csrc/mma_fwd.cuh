// MMA training forward: expected alignment (+ mass preservation) (+ expected soft attention),
// fused, one CTA per (batch*head) row, target-step recurrence kept in registers.
//
// Math (reference codebase/utils/monotonic_attention.py:40-69,
// 103-150, 183-193 and codebase/utils/functions.py:20-66), per row, step i, source j:
//   L_j  = log((1 - p_ij) + eps)             Cl_j = log(1 + eps) + sum_{k<j} L_k
//   cp_j = exp(Cl_j)    c_j = clamp(cp_j, eps, 1)    P_j = p_ij * cp_j
//   u_j  = alpha_{i-1,j} / c_j    s_j = sum_{k<=j} u_k    alpha_ij = clamp(P_j * s_j, 0, 1)
//   soft: m = max_j E_j   e_j = exp(E_j - m) + eps   D_j = eps + sum_{k<=j} e_k
//         r_j = alpha'_ij / D_j   R_j = sum_{k>=j} r_k   beta_ij = clamp(e_j * R_j, 0, 1)
//   (chunkwise: the two sums run over windows of c frames instead.)
// alpha' is alpha after mass preservation, which only touches one column (`last`), so its
// contribution to R is added analytically and the row sum shares the barrier of the R scan.
#pragma once

#include "mma_common.cuh"

namespace simulst {

// FULL = no padding mask and S == THREADS*VPT: every element of every thread is a live
// source position, so all per-element validity predicates vanish at compile time.
template <int THREADS, int VPT, typename T, int MODE, bool FULL>
__global__ void __launch_bounds__(THREADS) mma_fwd_kernel(const MmaParams prm, const StagePlan plan) {
    constexpr int NW = THREADS / kWarp;
    constexpr bool SOFT = MODE != kModeHard;
    constexpr bool CHUNK = MODE == kModeSoftCk;
    constexpr int NS = kFwdStages;
    static_assert(VPT % 4 == 0, "VPT must be a multiple of 4");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    float* bcast = xraw + 2 * kXSlots * kXStride;          // small broadcast area (32 floats)
    unsigned char* stage0 = smem + plan.header_bytes();
    float* win = reinterpret_cast<float*>(stage0 + (size_t)NS * plan.rows * plan.row_bytes);
    (void)win;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    if (row_filtered_out(prm, n)) return;       // this row belongs to the call's other pass
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool want_d = prm.delays != nullptr;
    const float fill = (prm.flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const bool vec_out = FULL || prm.vec_out != 0;

    const int ld_p = prm.ld_p, ld_e = prm.ld_e, ld_a = prm.ld_alpha, ld_b = prm.ld_beta;     // row pitches (elements)
    const T* gp = reinterpret_cast<const T*>(prm.p) + (size_t)n * T_len * ld_p;
    const T* ge = SOFT ? reinterpret_cast<const T*>(prm.e) + (size_t)n * T_len * ld_e : nullptr;
    float* g_alpha = prm.alpha + (size_t)n * T_len * ld_a;
    float* g_beta = SOFT ? prm.beta + (size_t)n * T_len * ld_b : nullptr;

    Xchg xc(xraw);

    // ---- per-row constants: validity bits, padding mask, column rewritten by mass preservation
    unsigned in_bits = 0u, live_bits = 0u;
    int n_live = 0;
    if constexpr (!FULL) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = j0 + k;
            if (j < S) {
                in_bits |= 1u << k;
                const bool padded = prm.mask != nullptr && prm.mask[(size_t)n * S + j] != 0;
                if (!padded) { live_bits |= 1u << k; ++n_live; }
            }
        }
    }
    auto is_in = [&](int k) -> bool { return FULL ? true : ((in_bits >> k) & 1u) != 0u; };
    auto is_live = [&](int k) -> bool { return FULL ? true : ((live_bits >> k) & 1u) != 0u; };
    // mass_preservation: no mask / left padding -> REPLACE column S-1 with the residual of
    // the other columns; right padding -> ADD the residual of all columns at src_len-1.
    const bool mp_add = !FULL && prm.mask != nullptr && !(prm.flags & SIMULST_MMA_LEFT_PADDING);
    int last = S - 1;
    const bool last_thread = tid == THREADS - 1;
    // does element k of this thread sit on the column rewritten by mass preservation?
    auto at_last = [&](int k) -> bool { return FULL ? (k == VPT - 1 && last_thread) : (j0 + k == last); };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (mp_add) {
        float cnt = warp_sum((float)n_live);
        if (lane == 0) xc.slot(0)[warp] = cnt;
    }
    __syncthreads();
    if (mp_add) {
        last = (int)combine_sum<NW>(xc.slot(0), lane) - 1;
        xc.flip();
    }

    // ---- row staging ring (TMA bulk copies when rows are 16-byte aligned)
    const unsigned row_bytes = (unsigned)(S * sizeof(T));
    auto stage_p = [&](int s) { return reinterpret_cast<T*>(stage0 + (size_t)(s * plan.rows) * plan.row_bytes); };
    auto stage_e = [&](int s) { return reinterpret_cast<T*>(stage0 + (size_t)(s * plan.rows + 1) * plan.row_bytes); };
    const bool use_tma = prm.tma || prm.tma_shift;
    auto issue = [&](int i, int s) {
        if (prm.tma) {
            if (tid == 0) {
                mbar_expect_tx(&bars[s], SOFT ? 2u * row_bytes : row_bytes);
                tma_load_1d(stage_p(s), gp + (size_t)i * ld_p, row_bytes, &bars[s]);
                if (SOFT) tma_load_1d(stage_e(s), ge + (size_t)i * ld_e, row_bytes, &bars[s]);
            }
        } else if (prm.tma_shift) {
            if (tid == 0) {
                unsigned np = 0u, ne = 0u;
                const void* sp = tma_span(gp + (size_t)i * ld_p, row_bytes, np);
                const void* se = SOFT ? tma_span(ge + (size_t)i * ld_e, row_bytes, ne) : nullptr;
                mbar_expect_tx(&bars[s], np + ne);
                tma_load_1d(stage_p(s), sp, np, &bars[s]);
                if (SOFT) tma_load_1d(stage_e(s), se, ne, &bars[s]);
            }
        } else {
            T* dp = stage_p(s);
            const T* sp = gp + (size_t)i * ld_p;
#pragma unroll 1
            for (int j = tid; j < S; j += THREADS) dp[j] = sp[j];
            if (SOFT) {
                T* de = stage_e(s);
                const T* se = ge + (size_t)i * ld_e;
#pragma unroll 1
                for (int j = tid; j < S; j += THREADS) de[j] = se[j];
            }
        }
    };
#pragma unroll
    for (int i = 0; i < NS - 1; ++i)
        if (i < T_len) issue(i, i);
    if (!use_tma) __syncthreads();

    const float one_eps = 1.0f + eps;       // first element of the exclusive cumprod (functions.py:28-33)
    float a_prev[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) a_prev[k] = (j0 + k == 0) ? 1.0f : 0.0f;
    bool bad = false;                       // some p outside [0,1] or NaN: classified at the end
    bool nan_out = false;

    int s = 0;                              // ring slot of step i
    unsigned parity = 0u;
    int s_fill = NS - 1;                    // ring slot refilled at step i (row i+NS-1)
    for (int i = 0; i < T_len; ++i) {
        // refill the slot consumed in the previous step (every thread passed >= 1 barrier since)
        if (i + NS - 1 < T_len) issue(i + NS - 1, s_fill);
        if (use_tma) mbar_wait(&bars[s], parity);

        float p[VPT], E[VPT];
        if (prm.tma_shift) {
            lds_row_shift<T, VPT>(stage_p(s), row_shift(gp + (size_t)i * ld_p), j0, p);
            if (SOFT) lds_row_shift<T, VPT>(stage_e(s), row_shift(ge + (size_t)i * ld_e), j0, E);
        } else {
            lds_row<T, VPT>(stage_p(s), j0, p);
            if (SOFT) lds_row<T, VPT>(stage_e(s), j0, E);
        }
        if (++s == NS) { s = 0; parity ^= 1u; }
        if (++s_fill == NS) s_fill = 0;

        // ---------------- step-invariant part: exclusive cumprod of (1-p)+eps, max of E
        float cpre[VPT];            // local exclusive product prefix
        float xtot = 1.0f, Emax = -INFINITY;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            if (is_in(k)) bad = bad || !(p[k] >= -1e-10f) || !(p[k] <= 1.0f);
            if (!is_live(k)) p[k] = 0.f;
            const float x = is_in(k) ? (1.0f - p[k]) + eps : 1.0f;
            cpre[k] = xtot;
            xtot *= x;
            if (SOFT) {
                if (!is_live(k)) E[k] = fill;
                if (is_in(k)) Emax = fmaxf(Emax, E[k]);
            }
        }
        const float xinc = warp_incl_prefix_mul(xtot, lane);
        if (lane == 31) xc.slot(0)[warp] = xinc;
        if (SOFT) {
            const float wm = warp_max(Emax);
            if (lane == 0) xc.slot(1)[warp] = wm;
        }
        const float xexc = lane_prev(xinc, lane, 1.0f);
        __syncthreads();
        const float xoff = combine_prefix_mul<NW>(xc.slot(0), warp, lane);
        float m = 0.f;
        if (SOFT) m = combine_max<NW>(xc.slot(1), lane);
        xc.flip();

        const float cbase = (one_eps * xoff) * xexc;
        float rc[VPT], P[VPT];      // 1/clamp(cp, eps, 1) and p*cp
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const float cp = cbase * cpre[k];
            rc[k] = fast_rcp(fminf(fmaxf(cp, eps), 1.0f));
            P[k] = p[k] * cp;
        }

        float ex[VPT], rD[VPT];     // e_j and 1/D_j
        float D_last = 0.f;
        if (SOFT) {
            float etot = 0.f;
            float Dl[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                ex[k] = is_in(k) ? (fast_exp(E[k] - m) + eps) : 0.f;
                etot += ex[k];
                Dl[k] = etot;            // local inclusive prefix
            }
            if (!CHUNK) {
                const float einc = warp_incl_prefix(etot, lane);
                if (lane == 31) xc.slot(0)[warp] = einc;
                const float eexc = lane_prev(einc, lane, 0.f);
                __syncthreads();
                const float2 eo = combine_prefix<NW>(xc.slot(0), warp, lane);
                xc.flip();
                const float ebase = eo.x + eexc;
#pragma unroll
                for (int k = 0; k < VPT; ++k) Dl[k] = eps + (ebase + Dl[k]);
                if (FULL) D_last = eps + eo.y;
                nan_out = nan_out || (etot != etot);
            } else {
                // D_j = eps + sum_{k=j-c+1..j} e_k  (moving_sum(e, c, 1))
#pragma unroll
                for (int k = 0; k < VPT; ++k) win[j0 + k] = ex[k];
                __syncthreads();
                const int cw = prm.chunk;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int j = j0 + k;
                    float acc = 0.f;
                    for (int q = max(0, j - cw + 1); q <= j; ++q) acc += win[q];
                    Dl[k] = eps + acc;
                }
            }
            // the thread that owns column `last` publishes D_last for the residual term
            if (mp && !(FULL && !CHUNK)) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (at_last(k)) bcast[0] = Dl[k];
            }
#pragma unroll
            for (int k = 0; k < VPT; ++k) rD[k] = fast_rcp(Dl[k]);
        }

        // ---------------- recurrence: alpha_i from alpha_{i-1}
        float utot = 0.f;
        float sloc[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            utot += a_prev[k] * rc[k];
            sloc[k] = utot;
        }
        const float uinc = warp_incl_prefix(utot, lane);
        if (lane == 31) xc.slot(0)[warp] = uinc;
        const float uexc = lane_prev(uinc, lane, 0.f);
        __syncthreads();        // also orders win[] reads (chunk) and bcast[0]
        const float2 uo = combine_prefix<NW>(xc.slot(0), warp, lane);
        xc.flip();
        if (SOFT && mp && !(FULL && !CHUNK)) D_last = bcast[0];
        const float ubase = uo.x + uexc;
        float a[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const float z = P[k] * (ubase + sloc[k]);
            a[k] = fminf(fmaxf(z, 0.0f), 1.0f);
            a_prev[k] = a[k];
        }
        nan_out = nan_out || (utot != utot);     // NaN anywhere in u poisons the thread total

        // ---------------- mass preservation + soft attention
        if (mp || SOFT || want_d) {
            float rsum = 0.f;           // row sum entering the residual
            float wloc = 0.f;           // sum of (j+1)*alpha over this thread's columns (expected delay)
            float rloc[VPT];            // r_j, then local inclusive suffix sums
            float a_last = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const bool is_last = at_last(k);
                if (is_last) a_last = a[k];
                if (want_d) wloc = __fmaf_rn(a[k], (float)(j0 + k + 1), wloc);
                const bool counted = mp_add || !is_last;
                if (mp && counted) rsum += a[k];
                if (SOFT) rloc[k] = ((mp && !counted) || !is_in(k)) ? 0.f : a[k] * rD[k];
            }
            float rinc = 0.f, rexc = 0.f;
            if (SOFT && !CHUNK) {
#pragma unroll
                for (int k = VPT - 2; k >= 0; --k) rloc[k] += rloc[k + 1];
                rinc = warp_incl_suffix(rloc[0], lane);
                if (lane == 0) xc.slot(0)[warp] = rinc;
                rexc = lane_next(rinc, lane, 0.f);
            }
            if (SOFT && CHUNK) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) win[j0 + k] = rloc[k];
            }
            if (mp) {
                const float ws = warp_sum(rsum);
                if (lane == 0) xc.slot(1)[warp] = ws;
            }
            if (want_d) {
                const float wd = warp_sum(wloc);
                if (lane == 0) xc.slot(2)[warp] = wd;
            }
            __syncthreads();
            float resid = 0.f, row_total = 0.f;
            if (mp) {
                row_total = combine_sum<NW>(xc.slot(1), lane);
                resid = 1.0f - fminf(fmaxf(row_total, 0.0f), 1.0f);
            }
            if (want_d) {
                // expected delay of the OUTPUT row (mma_criterion.py:146-157): the raw weighted
                // sum corrected for the column mass preservation rewrites (owner thread knows it)
                const float wtot = combine_sum<NW>(xc.slot(2), lane);
                float* dst = prm.delays + (size_t)n * T_len + i;
                if (!mp) {
                    if (tid == 0) *dst = wtot;
                } else {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (at_last(k)) *dst = wtot + (float)(last + 1) * (mp_add ? resid : (resid - a_last));
                }
            }
            float R[VPT];
            if (SOFT && !CHUNK) {
                const float2 ro = combine_suffix<NW>(xc.slot(0), warp, lane);
                const float rbase = ro.x + rexc;
                const float extra = mp ? resid * fast_rcp(D_last) : 0.f;
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    R[k] = (rbase + rloc[k]) + ((mp && (FULL || j0 + k <= last)) ? extra : 0.f);
            }
            if (SOFT && CHUNK) {
                const int cw = prm.chunk;
                const float extra = mp ? resid * fast_rcp(D_last) : 0.f;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int j = j0 + k;
                    float acc = 0.f;
                    for (int q = j; q <= min(S - 1, j + cw - 1); ++q) acc += win[q];
                    if (mp && j <= last && last <= j + cw - 1) acc += extra;
                    R[k] = acc;
                }
            }
            xc.flip();
            // outputs
            if (mp) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (at_last(k)) {
                        a[k] = mp_add ? (a[k] + resid) : resid;
                        if (prm.side != nullptr) {
                            float* sd = prm.side + ((size_t)n * T_len + i) * 2;
                            sd[0] = a_last;
                            sd[1] = row_total;
                        }
                    }
            }
            if (SOFT) {
                float b[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float v = is_live(k) ? ex[k] * R[k] : 0.f;
                    b[k] = fminf(fmaxf(v, 0.0f), 1.0f);
                }
                st_row_f32<VPT, FULL>(g_beta + (size_t)i * ld_b, j0, S, vec_out, b);
            }
        }
        st_row_f32<VPT, FULL>(g_alpha + (size_t)i * ld_a, j0, S, vec_out, a);
    }

    // ---- data-error reporting (prob_check / safe_cumprod semantics), slow path only on error
    if (prm.status != nullptr) {
        if (nan_out) atomicOr(prm.status, SIMULST_ST_NAN);
        if (bad) {
            unsigned bits = 0u;
            for (int i = 0; i < T_len; ++i)
                for (int k = 0; k < VPT; ++k)
                    if (j0 + k < S) {
                        const float v = to_f32<T>(gp[(size_t)i * ld_p + j0 + k]);
                        bits |= prob_bits(v);
                        if ((1.0f - v) + eps < 0.f) bits |= SIMULST_ST_NEGPROD;
                    }
            atomicOr(prm.status, bits);
        }
    }
}

// ------------------------------------------------------------------ host-side launcher
template <int THREADS, int VPT, typename T, int MODE, bool FULL>
int launch_mma_fwd_impl(const MmaParams& prm, cudaStream_t stream) {
    StagePlan plan;
    plan.rows = (MODE == kModeHard) ? 1 : 2;
    plan.row_bytes = ((THREADS * VPT * (int)sizeof(T)) + 16 + 127) / 128 * 128;    // + the head of a shifted row
    plan.win_floats = (MODE == kModeSoftCk) ? THREADS * VPT : 0;
    plan.n_stage = kFwdStages;
    auto kern = mma_fwd_kernel<THREADS, VPT, T, MODE, FULL>;
    static bool attr_done[64] = {};     // per device; only ever flips false -> true
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total()) != cudaSuccess) {
            cudaGetLastError();
            return SIMULST_E_LAUNCH;
        }
        attr_done[dev & 63] = true;
    }
    kern<<<prm.N, THREADS, plan.total(), stream>>>(prm, plan);
    return check_launch();
}

template <int THREADS, int VPT, typename T, int MODE>
int launch_mma_fwd(const MmaParams& prm, cudaStream_t stream) {
    const bool full = prm.mask == nullptr && prm.S == THREADS * VPT && prm.vec_out && prm.tma;
    return full ? launch_mma_fwd_impl<THREADS, VPT, T, MODE, true>(prm, stream)
                : launch_mma_fwd_impl<THREADS, VPT, T, MODE, false>(prm, stream);
}

}  // namespace simulst
