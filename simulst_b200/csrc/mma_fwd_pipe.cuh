// MMA training forward, software-pipelined: expected alignment (+ mass preservation)
// (+ infinite-lookback expected soft attention), one CTA per (batch*head) row.
//
// Per target step i the work splits into a step-INVARIANT part (functions of p_i / E_i only:
// exclusive cumprod, clamped divisor, P, exp, the D prefix) and the RECURRENCE
// (u-prefix over alpha_{i-1} -> alpha_i -> r-suffix -> beta_i).  Loop iteration i runs the
// recurrence of step i and the invariant part of step i+1 in the same instruction stream and
// lets their block-wide scans share barriers: 2 __syncthreads per step instead of 4, and two
// independent dependency chains for the scheduler to interleave.
//
//   phase A   rec: u-prefix (local + warp)        inv: load row i+1, cumprod (local + warp), max E
//   ---- barrier A  (then: TMA refill of the ring slot just read)
//   phase B   rec: alpha_i, r-suffix, row sum     inv: cp, 1/c, P, exp, e-prefix (local + warp)
//   ---- barrier B
//   phase C   rec: beta_i, stores                 inv: 1/D
//
// Rows of p_choose / soft_energy arrive through a ring of TMA 1-D bulk copies (UBLKCP) issued
// NS steps ahead by one thread.  Math / reference lines: see mma_steps.cuh and mma_fwd.cuh.
// Chunkwise soft attention and rows that TMA cannot stage (not 16-byte aligned) use the
// generic kernel in mma_fwd.cuh.
#pragma once

#include "mma_common.cuh"
#include "mma_steps.cuh"

namespace simulst {

constexpr int kPipeStages = 4;          // deepest row-staging ring (shallower when rows are long)
constexpr int kPipeMaxThreads = 512;    // 1024-thread CTAs (64 registers/thread) keep the generic kernel

template <int THREADS, int VPT, typename T, bool SOFT, bool FULL>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 4 : (THREADS <= 256 ? 2 : 1)))
mma_fwd_pipe_kernel(const MmaParams prm, const StagePlan plan) {
    constexpr int NW = THREADS / kWarp;
    constexpr int H = VPT / 2;
    static_assert(VPT % 4 == 0, "VPT must be a multiple of 4");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xA = reinterpret_cast<float*>(smem + 128);       // exchange buffer of barrier A
    float* xB = xA + kXSlots * kXStride;                    // exchange buffer of barrier B
    float* bcast = xA + 2 * kXSlots * kXStride;             // [2] 1/D at the mass-preservation column
    unsigned char* stage0 = smem + plan.header_bytes();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const float fill = (prm.flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const bool vec_out = FULL || prm.vec_out != 0;
    const int NS = plan.n_stage;

    const T* gp = reinterpret_cast<const T*>(prm.p) + (size_t)n * T_len * S;
    const T* ge = SOFT ? reinterpret_cast<const T*>(prm.e) + (size_t)n * T_len * S : nullptr;
    float* g_alpha = prm.alpha + (size_t)n * T_len * S;
    float* g_beta = SOFT ? prm.beta + (size_t)n * T_len * S : nullptr;

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
    // mass_preservation: no mask / left padding -> REPLACE column S-1 with the residual of the
    // other columns; right padding -> ADD the residual of all columns at src_len-1.
    const bool mp_add = !FULL && prm.mask != nullptr && !(prm.flags & SIMULST_MMA_LEFT_PADDING);
    int last = S - 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (mp_add) {
        const float cnt = warp_sum((float)n_live);
        if (lane == 0) xA[warp] = cnt;
    }
    __syncthreads();
    if (mp_add) {
        last = (int)combine_sum<NW>(xA, lane) - 1;
        __syncthreads();
    }
    // element of this thread that sits on the mass-preservation column (-1: none)
    int k_last = -1;
    if (FULL) {
        if (tid == THREADS - 1) k_last = VPT - 1;
    } else if (last >= j0 && last < j0 + VPT) {
        k_last = last - j0;
    }
    const bool own_last = mp && k_last >= 0;

    // ---- row staging ring
    const unsigned row_bytes = (unsigned)(S * sizeof(T));
    auto stage_p = [&](int s) { return reinterpret_cast<T*>(stage0 + (size_t)(s * plan.rows) * plan.row_bytes); };
    auto stage_e = [&](int s) { return reinterpret_cast<T*>(stage0 + (size_t)(s * plan.rows + 1) * plan.row_bytes); };
    auto issue = [&](int i, int s) {      // called by thread 0 only
        mbar_expect_tx(&bars[s], SOFT ? 2u * row_bytes : row_bytes);
        tma_load_1d(stage_p(s), gp + (size_t)i * S, row_bytes, &bars[s]);
        if (SOFT) tma_load_1d(stage_e(s), ge + (size_t)i * S, row_bytes, &bars[s]);
    };
    if (tid == 0) {
        for (int i = 0; i < NS && i < T_len; ++i) issue(i, i);
    }

    const float one_eps = 1.0f + eps;       // first element of the exclusive cumprod (functions.py:28-33)
    // recurrence state and the invariants of the step the recurrence is about to run
    float2 a_prev[H], rc[H], P[H], ex[H], rD[H];
#pragma unroll
    for (int q = 0; q < H; ++q) {
        a_prev[q] = make_float2((j0 + 2 * q == 0) ? 1.0f : 0.0f, 0.0f);
        rc[q] = P[q] = ex[q] = rD[q] = f2(0.f);
    }
    unsigned umax = 0u;                     // FULL rows: first-level prob_check
    bool bad = false;                       // ragged rows: exact per-element check
    bool nan_out = false;

    int slot = 0;                           // ring slot holding the row of the NEXT invariant step
    unsigned parity = 0u;

    // One loop iteration: recurrence of step `i` (REC) + invariant part of step `i + 1` (INV).
    auto body = [&](auto rec_c, auto inv_c, const int i) {
        constexpr bool REC = decltype(rec_c)::value;
        constexpr bool INV = decltype(inv_c)::value;

        // ================================================= phase A
        float2 sl[H];
        float uinc = 0.f, uexc = 0.f;
        if constexpr (REC) {
            const float ut = local_u_prefix<VPT>(a_prev, rc, sl);
            nan_out = nan_out || (ut != ut);     // NaN anywhere in u poisons the thread total
            uinc = wscan_prefix_add(ut);
            if (lane == 31) xA[0 * kXStride + warp] = uinc;
            uexc = wprev(uinc, 0.f);
        }
        float2 p_n[H], E_n[H], cpre[H];
        float xinc = 1.f, xexc = 1.f;
        if constexpr (INV) {
            mbar_wait(&bars[slot], parity);
            lds_row2<T, VPT, FULL>(stage_p(slot) + j0, p_n, umax);
            if (SOFT) {
                unsigned dummy = 0u;
                lds_row2<T, VPT, false>(stage_e(slot) + j0, E_n, dummy);
            }
            if constexpr (!FULL) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    float& pk = SIMULST_EL(p_n, k);
                    if (is_in(k)) bad = bad || !(pk >= -1e-10f) || !(pk <= 1.0f);
                    if (!is_live(k)) pk = 0.f;
                    if (SOFT) {
                        float& ek = SIMULST_EL(E_n, k);
                        if (!is_live(k)) ek = is_in(k) ? fill : -INFINITY;
                    }
                }
            }
            const float xt = local_cumprod<VPT>(p_n, eps, cpre);
            xinc = wscan_prefix_mul(xt);
            if (lane == 31) xA[1 * kXStride + warp] = xinc;
            xexc = wprev(xinc, 1.f);
            if (SOFT) {
                float em = fmaxf(E_n[0].x, E_n[0].y);
#pragma unroll
                for (int q = 1; q < H; ++q) em = fmaxf(em, fmaxf(E_n[q].x, E_n[q].y));
                const float wm = wmax_redux(em);
                if (lane == 0) xA[2 * kXStride + warp] = wm;
            }
        }
        __syncthreads();                    // ---- barrier A
        if constexpr (INV) {
            // every thread has read ring slot `slot`: refill it with the row NS steps ahead
            if (tid == 0 && i + 1 + NS < T_len) issue(i + 1 + NS, slot);
            if (++slot == NS) { slot = 0; parity ^= 1u; }
        }

        // ================================================= phase B
        float2 Rl[H];
        float rexc = 0.f, a_last_raw = 0.f;
        if constexpr (REC) {
            const float ubase = xw_prefix_add<NW>(xA + 0 * kXStride, warp, lane) + uexc;
            float2 sfull[H], z[H];
            finish_u_prefix<VPT>(ubase, sl, P, sfull, z);
#pragma unroll
            for (int q = 0; q < H; ++q) a_prev[q] = min2(z[q], 1.0f);      // z >= 0: P >= 0, s >= 0
            if (mp || SOFT) {
                // alpha entering the row sum / the soft-attention numerator: the mass-preservation
                // column is left out when it is REPLACED (its residual is added analytically)
                float2 a_s[H];
#pragma unroll
                for (int q = 0; q < H; ++q) a_s[q] = a_prev[q];
                if (own_last) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if ((FULL ? k == VPT - 1 : true) && k == k_last) {
                            a_last_raw = SIMULST_EL(a_s, k);
                            if (!mp_add) SIMULST_EL(a_s, k) = 0.f;
                        }
                }
                if (SOFT) {
                    const float rt = local_r_suffix<VPT>(a_s, rD, Rl);
                    const float rinc = wscan_suffix_add(rt);
                    if (lane == 0) xB[0 * kXStride + warp] = rinc;
                    rexc = wnext(rinc, 0.f);
                }
                if (mp) {
                    float2 acc = a_s[0];
#pragma unroll
                    for (int q = 1; q < H; ++q) acc = add2(acc, a_s[q]);
                    const float ws = warp_sum(acc.x + acc.y);
                    if (lane == 0) xB[1 * kXStride + warp] = ws;
                }
            }
        }
        float2 ex_n[H], Dl[H], rc_n[H], P_n[H];
        float eexc = 0.f;
        if constexpr (INV) {
            const float xoff = xw_prefix_mul<NW>(xA + 1 * kXStride, warp, lane);
            const float cbase = (one_eps * xoff) * xexc;
            float2 cp[H];
            finish_cumprod<VPT>(cbase, cpre, p_n, eps, cp, rc_n, P_n);
            if (SOFT) {
                const float m = xw_max<NW>(xA + 2 * kXStride, lane);
                float2 unused[H];
                const float et = local_exp_prefix<VPT, false>(E_n, m, eps, unused, ex_n, Dl);
                nan_out = nan_out || (et != et);
                const float einc = wscan_prefix_add(et);
                if (lane == 31) xB[2 * kXStride + warp] = einc;
                eexc = wprev(einc, 0.f);
            }
        }
        __syncthreads();                    // ---- barrier B

        // ================================================= phase C
        if constexpr (REC) {
            float resid = 0.f, row_total = 0.f;
            if (mp) {
                row_total = xw_sum<NW>(xB + 1 * kXStride, lane);
                resid = 1.0f - fminf(fmaxf(row_total, 0.0f), 1.0f);
            }
            if (SOFT) {
                // R_k = sum_{q>=k} r_q  (+ resid / D_last for every k <= last; positions beyond
                // `last` are padded or outside the row, their beta is zero anyway)
                float rbase = xw_suffix_add<NW>(xB + 0 * kXStride, warp, lane) + rexc;
                if (mp) rbase += resid * bcast[i & 1];
                const float2 rb = f2(rbase);
                float2 b[H];
#pragma unroll
                for (int q = 0; q < H; ++q) b[q] = min2(mul2(ex[q], add2(rb, Rl[q])), 1.0f);
                if constexpr (!FULL) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (!is_live(k)) SIMULST_EL(b, k) = 0.f;
                }
                st_row2_f32<VPT, FULL>(g_beta + (size_t)i * S, j0, S, vec_out, b);
            }
            if (own_last) {
                float2 a_out[H];
#pragma unroll
                for (int q = 0; q < H; ++q) a_out[q] = a_prev[q];
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if ((FULL ? k == VPT - 1 : true) && k == k_last)
                        SIMULST_EL(a_out, k) = mp_add ? (a_last_raw + resid) : resid;
                st_row2_f32<VPT, FULL>(g_alpha + (size_t)i * S, j0, S, vec_out, a_out);
                if (prm.side != nullptr)
                    *reinterpret_cast<float2*>(prm.side + ((size_t)n * T_len + i) * 2) = make_float2(a_last_raw, row_total);
            } else {
                st_row2_f32<VPT, FULL>(g_alpha + (size_t)i * S, j0, S, vec_out, a_prev);
            }
        }
        if constexpr (INV) {
#pragma unroll
            for (int q = 0; q < H; ++q) { rc[q] = rc_n[q]; P[q] = P_n[q]; }
            if (SOFT) {
                const float ebase = xw_prefix_add<NW>(xB + 2 * kXStride, warp, lane) + eexc;
                finish_exp_prefix<VPT>(ebase, eps, Dl, rD);
#pragma unroll
                for (int q = 0; q < H; ++q) ex[q] = ex_n[q];
                if (own_last) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if ((FULL ? k == VPT - 1 : true) && k == k_last) bcast[(i + 1) & 1] = SIMULST_EL(rD, k);
                }
            }
        }
    };

    using Yes = std::integral_constant<bool, true>;
    using No = std::integral_constant<bool, false>;
    body(No{}, Yes{}, -1);                                   // prologue: invariants of step 0
    for (int i = 0; i < T_len - 1; ++i) body(Yes{}, Yes{}, i);
    body(Yes{}, No{}, T_len - 1);                            // epilogue: last recurrence step

    // ---- data-error reporting (prob_check / safe_cumprod semantics), slow path only on error
    if (prm.status != nullptr) {
        if (nan_out) atomicOr(prm.status, SIMULST_ST_NAN);
        if (FULL) bad = umax_trips<T>(umax);
        if (bad) {
            unsigned bits = 0u;
            for (int i = 0; i < T_len; ++i)
                for (int k = 0; k < VPT; ++k)
                    if (j0 + k < S) {
                        const float v = to_f32<T>(gp[(size_t)i * S + j0 + k]);
                        bits |= prob_bits(v);
                        if ((1.0f - v) + eps < 0.f) bits |= SIMULST_ST_NEGPROD;
                    }
            if (bits) atomicOr(prm.status, bits);
        }
    }
}

// ------------------------------------------------------------------ host-side launcher
template <int THREADS, int VPT, typename T, bool SOFT, bool FULL>
int launch_mma_fwd_pipe_impl(const MmaParams& prm, cudaStream_t stream) {
    StagePlan plan;
    plan.rows = SOFT ? 2 : 1;
    plan.row_bytes = ((THREADS * VPT * (int)sizeof(T)) + 127) / 128 * 128;
    plan.win_floats = 0;
    plan.n_stage = kPipeStages;
    // keep 4 CTAs of a 128-thread configuration resident (<= 56 KB each); long rows: what fits
    const size_t budget = THREADS <= 128 ? 56 * 1024 : (THREADS <= 256 ? 110 * 1024 : 220 * 1024);
    while (plan.n_stage > 1 && plan.total() > budget) --plan.n_stage;
    auto kern = mma_fwd_pipe_kernel<THREADS, VPT, T, SOFT, FULL>;
    static size_t attr_set[64] = {};    // per device: largest dynamic smem size opted into
    int dev = 0;
    cudaGetDevice(&dev);
    if (plan.total() > attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total()) != cudaSuccess) {
            cudaGetLastError();
            return SIMULST_E_LAUNCH;
        }
        attr_set[dev & 63] = plan.total();
    }
    kern<<<prm.N, THREADS, plan.total(), stream>>>(prm, plan);
    return check_launch();
}

// SOFT here means infinite lookback; requires prm.tma (16-byte aligned rows).
template <int THREADS, int VPT, typename T, bool SOFT>
int launch_mma_fwd_pipe(const MmaParams& prm, cudaStream_t stream) {
    const bool full = prm.mask == nullptr && prm.S == THREADS * VPT && prm.vec_out;
    return full ? launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true>(prm, stream)
                : launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, false>(prm, stream);
}

}  // namespace simulst
