// MMA training backward, software-pipelined: gradients of (expected alignment -> mass
// preservation -> infinite-lookback expected soft attention) w.r.t. p_choose and soft_energy.
// One CTA per (batch*head) row walks the target steps in reverse; every scan of the forward
// pass is recomputed (shared code: mma_steps.cuh), only the forward OUTPUT alpha and the
// [T,2] side vector are read back.
//
// Derivation: SURVEY Appendix A.2-A.4.  Per step i the work splits into
//   EARLY(i)  everything that does not depend on the recurrence gradient ("carry"):
//             S1  cp = excl-cumprod(1-p), rc = 1/clamp(cp), P = p*cp ; e = exp(E-m)+eps, 1/D
//             S2  s = prefix(alpha_{i-1}*rc) ; R = suffix(r), r = alpha'_i/D ; W = suffix(r/D)
//             S3  gr = prefix(gR), gR = gb*e, gb = gbeta*1[b<=1] ; V = suffix(gR*W)
//                 -> d/dalpha' = gr/D ;  suffix(gD) = -(gr_excl*W + V)   (gD = -gr*r/D: the
//                 second-level scan is expanded so that it shares the barrier of gr)
//                 -> gE = (gb*R + suffix(gD))*(e-eps) - [argmax]*sum(.)
//   LATE(i)   S5  g0 = g'' + carry ; gu = suffix(mz*P*g0) ; carry' = gu*rc
//             S6  gL = excl-suffix(g0*c1 - gu*c2) ; gp = g0*c3 - gL/((1-p)+eps)
//             with c1 = mz*s*p*cp, c2 = rc*u*1[eps<=cp<=1]*cp, c3 = mz*s*cp prepared by EARLY.
// Loop iteration `s` runs MAXS(s-1) (row max / arg-max of the energies), EARLY(s) and
// LATE(s+1) in one instruction stream; their block-wide scans share THREE barriers
// (two for hard attention) instead of the generic kernel's six:
//   B1: x-prefix-product, e-prefix, gu-suffix, sum(gE) of the previous step, max/arg-max
//   B2: u-prefix, R-suffix, W-suffix, gL-suffix
//   B3: gr-prefix, V-suffix
// EARLY hands {g'', mz*P, rc, c2, c3} to LATE through a thread-private shared-memory stash.
// Every global read is a TMA 1-D bulk copy (UBLKCP) issued a step ahead by one thread: rings
// for p / energy / alpha (3 deep: a row is visited by two consecutive iterations), a single
// slot for grad_beta, and grad_alpha lands directly in the (double-buffered) g'' slot of the
// stash, where EARLY turns it into g'' in place.
#pragma once

#include "mma_common.cuh"
#include "mma_steps.cuh"

namespace simulst {

constexpr int kBwdSlots = 12;   // exchange slots: B1 0..5, B2 6..9, B3 10..11

struct BwdPipePlan {
    int row_t_bytes;   // bytes reserved per staged p / energy row (multiple of 128)
    int row_f_bytes;   // bytes reserved per staged alpha row
    int soft;
    int stash_bytes;   // 4 arrays * THREADS * VPT * 4 (mz*P, rc, c2, c3)
    __host__ __device__ int header_bytes() const { return 128 + kBwdSlots * kXStride * 4 + 128; }
    // p ring (3) [+ energy ring (3)] + alpha ring (3) + g'' slots (2) [+ grad_beta slot] + stash
    __host__ __device__ size_t total() const {
        return (size_t)header_bytes() + (size_t)(soft ? 6 : 3) * row_t_bytes +
               (size_t)(soft ? 6 : 5) * row_f_bytes + (size_t)stash_bytes;
    }
};

template <int THREADS, int VPT, typename T, bool SOFT, bool FULL>
__global__ void __launch_bounds__(THREADS, (THREADS * VPT <= 1024 ? 4 : (THREADS <= 256 ? 2 : 1)))
mma_bwd_pipe_kernel(const MmaParams prm, const BwdPipePlan plan) {
    constexpr int NW = THREADS / kWarp;
    constexpr int H = VPT / 2;
    constexpr int Q4 = VPT / 4;
    static_assert(VPT % 4 == 0, "VPT must be a multiple of 4");

    extern __shared__ __align__(128) unsigned char smem[];
    // mbarriers: [0..2] p ring, [3..5] energy ring, [6..8] alpha ring, [9] grad_beta, [10..11] grad_alpha
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xb = reinterpret_cast<float*>(smem + 128);         // [kBwdSlots][32]
    float* bcast = xb + kBwdSlots * kXStride;                 // [0] 1/D at the mass-preservation column
    unsigned char* ring_p = smem + plan.header_bytes();
    unsigned char* ring_e = ring_p + 3 * plan.row_t_bytes;
    unsigned char* ring_a = ring_e + (SOFT ? 3 * plan.row_t_bytes : 0);
    unsigned char* slot_g = ring_a + 3 * plan.row_f_bytes;                       // 2 x g'' / grad_alpha (linear rows)
    unsigned char* slot_b = slot_g + 2 * plan.row_f_bytes;                       // grad_beta row
    float4* stash = reinterpret_cast<float4*>(slot_b + (SOFT ? plan.row_f_bytes : 0));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const float fill = (prm.flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const bool vec = FULL || prm.vec_out != 0;
    const bool has_ga = prm.g_alpha != nullptr;
    const bool has_gb = SOFT && prm.g_beta != nullptr;

    const size_t row0 = (size_t)n * T_len * S;
    const T* gp_in = reinterpret_cast<const T*>(prm.p) + row0;
    const T* ge_in = SOFT ? reinterpret_cast<const T*>(prm.e) + row0 : nullptr;
    const float* al = prm.alpha + row0;
    const float* gA_in = has_ga ? prm.g_alpha + row0 : nullptr;
    const float* gB_in = has_gb ? prm.g_beta + row0 : nullptr;
    T* gp_out = reinterpret_cast<T*>(prm.g_p) + row0;
    T* ge_out = SOFT ? reinterpret_cast<T*>(prm.g_e) + row0 : nullptr;
    const float* side = mp ? prm.side + (size_t)n * T_len * 2 : nullptr;

    // ---- per-row constants
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
    const bool mp_add = !FULL && prm.mask != nullptr && !(prm.flags & SIMULST_MMA_LEFT_PADDING);
    int last = S - 1;

    if (tid == 0) {
        for (int b = 0; b < 12; ++b) mbar_init(&bars[b], 1);
        mbar_fence_init();
    }
    if (mp_add) {
        const float cnt = warp_sum((float)n_live);
        if (lane == 0) xb[warp] = cnt;
    }
    __syncthreads();
    if (mp_add) {
        last = (int)combine_sum<NW>(xb, lane) - 1;
        __syncthreads();
    }
    int k_last = -1;
    if (FULL) {
        if (tid == THREADS - 1) k_last = VPT - 1;
    } else if (last >= j0 && last < j0 + VPT) {
        k_last = last - j0;
    }
    const bool own_last = mp && k_last >= 0;
    auto at_last = [&](int k) -> bool { return (FULL ? k == VPT - 1 : true) && k == k_last; };

    // ---- rings.  Rows are consumed top-down; ring position of a row = its consumption order.
    const unsigned t_bytes = (unsigned)(S * sizeof(T)), f_bytes = (unsigned)(S * 4);
    auto slot_p = [&](int k3) { return reinterpret_cast<const T*>(ring_p + (size_t)k3 * plan.row_t_bytes); };
    auto slot_e = [&](int k3) { return reinterpret_cast<const T*>(ring_e + (size_t)k3 * plan.row_t_bytes); };
    auto slot_a = [&](int k3) { return reinterpret_cast<const float*>(ring_a + (size_t)k3 * plan.row_f_bytes); };
    // thread 0 only.  Row r of p / energy / alpha goes to ring position (T-1-r) % 3.
    auto issue_p = [&](int r) {
        const int k = (T_len - 1 - r) % 3;
        mbar_expect_tx(&bars[k], t_bytes);
        tma_load_1d(const_cast<T*>(slot_p(k)), gp_in + (size_t)r * S, t_bytes, &bars[k]);
    };
    auto issue_e = [&](int r) {
        const int k = (T_len - 1 - r) % 3;
        mbar_expect_tx(&bars[3 + k], t_bytes);
        tma_load_1d(const_cast<T*>(slot_e(k)), ge_in + (size_t)r * S, t_bytes, &bars[3 + k]);
    };
    auto issue_a = [&](int r) {
        const int k = (T_len - 1 - r) % 3;
        mbar_expect_tx(&bars[6 + k], f_bytes);
        tma_load_1d(const_cast<float*>(slot_a(k)), al + (size_t)r * S, f_bytes, &bars[6 + k]);
    };
    // grad_alpha row r -> g'' slot (T-1-r) & 1 ; grad_beta row r -> the single slot
    auto gslot = [&](int k2) { return reinterpret_cast<float*>(slot_g + (size_t)k2 * plan.row_f_bytes); };
    auto issue_ga = [&](int r) {
        const int k = (T_len - 1 - r) & 1;
        mbar_expect_tx(&bars[10 + k], f_bytes);
        tma_load_1d(gslot(k), gA_in + (size_t)r * S, f_bytes, &bars[10 + k]);
    };
    auto issue_gb = [&](int r) {
        mbar_expect_tx(&bars[9], f_bytes);
        tma_load_1d(slot_b, gB_in + (size_t)r * S, f_bytes, &bars[9]);
    };
    if (tid == 0) {
        if (has_ga) issue_ga(T_len - 1);
        if (has_gb) issue_gb(T_len - 1);
        issue_p(T_len - 1);
        if (SOFT) {
            issue_e(T_len - 1);
            if (T_len >= 2) issue_e(T_len - 2);
        }
        issue_a(T_len - 1);
        if (T_len >= 2) issue_a(T_len - 2);
    }

    const float one_eps = 1.0f + eps;
    // ---- state carried between iterations
    float2 carry[H];            // dL/d alpha_{s+1} flowing into LATE(s+1)
#pragma unroll
    for (int q = 0; q < H; ++q) carry[q] = f2(0.f);
    float m_cur = 0.f;          // row max / arg-max of step s (from MAXS one iteration earlier)
    int amax_cur = -1;
    float gEsum_prev = 0.f;     // thread-local sum of gE*(e-eps) of step s+1, reduced at B1
    float gEm_fix = 0.f;        // this thread's value at the arg-max column of step s+1 (if it owns it)
    int fix_col = -1;
    // consumption-order ring position / parity of row s (p, energy, alpha alike): (T-1-s) % 3;
    // row s-1 sits one position further, row s+1 one before
    int kp = 0;
    unsigned par3 = 0u;

    auto stash_ld = [&](int arr, float2 (&v)[H]) {
#pragma unroll
        for (int q = 0; q < Q4; ++q) {
            const float4 t = stash[(arr * Q4 + q) * THREADS + tid];
            v[2 * q] = f2(t.x, t.y);
            v[2 * q + 1] = f2(t.z, t.w);
        }
    };
    auto stash_st = [&](int arr, const float2 (&v)[H]) {
#pragma unroll
        for (int q = 0; q < Q4; ++q)
            stash[(arr * Q4 + q) * THREADS + tid] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
    };
    auto lin_ld = [&](const float* row, float2 (&v)[H]) {     // linear fp32 row in shared memory
        unsigned dummy = 0u;
        lds_row2<float, VPT, false>(row + j0, v, dummy);
    };
    auto lin_st = [&](float* row, const float2 (&v)[H]) {
#pragma unroll
        for (int q = 0; q < Q4; ++q)
            *reinterpret_cast<float4*>(row + j0 + 4 * q) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
    };

    auto body = [&](auto steady_c, const int s) __attribute__((always_inline)) {
        constexpr bool STEADY = decltype(steady_c)::value;
        const bool doM = SOFT && (STEADY || (s - 1 >= 0 && s - 1 < T_len));
        const bool doE = STEADY || (s >= 0 && s < T_len);
        const bool doL = STEADY || (s + 1 >= 0 && s + 1 < T_len);
        int km = kp + 1;                    // ring position / parity of row s-1
        unsigned parm = par3;
        if (km == 3) { km = 0; parm ^= 1u; }
        int kl = kp - 1;                    // ring position of row s+1
        if (kl < 0) kl = 2;
        const int kg = (T_len - 1 - s) & 1;             // g'' / grad_alpha slot of step s
        const unsigned parg = ((unsigned)(T_len - 1 - s) >> 1) & 1u;
        const unsigned parb = (unsigned)(T_len - 1 - s) & 1u;

        // ================================================================ PRE-B1
        // ---- MAXS(s-1): row max and first arg-max of the energies
        float wm = -INFINITY;
        int wcand = 0x7fffffff;
        if (SOFT && doM) {
            mbar_wait(&bars[3 + km], parm);
            float2 Em[H];
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(slot_e(km) + j0, Em, dummy);
            if constexpr (!FULL) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (!is_live(k)) SIMULST_EL(Em, k) = is_in(k) ? fill : -INFINITY;
            }
            float em = fmaxf(Em[0].x, Em[0].y);
#pragma unroll
            for (int q = 1; q < H; ++q) em = fmaxf(em, fmaxf(Em[q].x, Em[q].y));
            wm = wmax_redux(em);
            int cand = 0x7fffffff;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k)
                if (is_in(k) && SIMULST_EL(Em, k) == wm) cand = j0 + k;
            wcand = __reduce_min_sync(kFull, cand);
        }
        // ---- EARLY(s) S1: thread-local cumprod and exp prefix
        float2 p_s[H], cpre[H], ex[H], Dl[H];
        float xinc = 1.f, einc = 0.f;
        if (doE) {
            mbar_wait(&bars[kp], par3);
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(slot_p(kp) + j0, p_s, dummy);
            float2 E_s[H];
            if (SOFT) lds_row2<T, VPT, false>(slot_e(kp) + j0, E_s, dummy);
            if constexpr (!FULL) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    if (!is_live(k)) SIMULST_EL(p_s, k) = 0.f;
                    if (SOFT) {
                        if (!is_live(k)) SIMULST_EL(E_s, k) = is_in(k) ? fill : -INFINITY;
                    }
                }
            }
            xinc = local_cumprod<VPT>(p_s, eps, cpre);
            if (SOFT) {
                float2 unused[H];
                einc = local_exp_prefix<VPT, false>(E_s, m_cur, eps, unused, ex, Dl);
            }
        }
        // ---- LATE(s+1) S5: g0 = g'' + carry, thread-local suffix of mz*P*g0
        float2 g0[H], Al[H];
        float ginc = 0.f;
        if (doL) {
            float2 gpp[H], Pm[H];
            lin_ld(gslot(kg ^ 1), gpp);          // g''(s+1)
            stash_ld(0, Pm);
#pragma unroll
            for (int q = 0; q < H; ++q) g0[q] = add2(gpp[q], carry[q]);
            float at = 0.f;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k) {
                at = fmaf(SIMULST_EL(Pm, k), SIMULST_EL(g0, k), at);
                SIMULST_EL(Al, k) = at;
            }
            ginc = at;
        }
        // ---- warp level
        float qinc = gEsum_prev;
        float xexc, eexc = 0.f, gexc;
        if (SOFT) {
            wscan_b1(xinc, einc, qinc, ginc);
            wneigh_b1(xinc, einc, ginc, xexc, eexc, gexc);
            if (lane == 31) xb[1 * kXStride + warp] = einc;
            if (lane == 31) xb[3 * kXStride + warp] = qinc;
            if (lane == 0) xb[4 * kXStride + warp] = wm;
            if (lane == 0) reinterpret_cast<int*>(xb)[5 * kXStride + warp] = wcand;
        } else {
            wscan_xd(xinc, ginc);
            xexc = wprev(xinc, 1.f);
            gexc = wnext(ginc, 0.f);
        }
        if (lane == 31) xb[0 * kXStride + warp] = xinc;
        if (lane == 0) xb[2 * kXStride + warp] = ginc;

        __syncthreads();                    // ================================ B1

        // ---- MAXS(s-1) result (becomes m_cur / amax_cur at the end of the iteration)
        float m_nxt = 0.f;
        int amax_nxt = -1;
        if (SOFT && doM) {
            m_nxt = xw_max<NW>(xb + 4 * kXStride, lane);
            const int* ci = reinterpret_cast<const int*>(xb) + 5 * kXStride;
            amax_nxt = 0x7fffffff;
#pragma unroll
            for (int w = 0; w < NW; ++w)
                if (xb[4 * kXStride + w] == m_nxt) amax_nxt = min(amax_nxt, ci[w]);
        }
        // ---- deferred arg-max correction of gE of step s+1 (its row was stored last iteration)
        if (SOFT && doL) {
            const float gEall = xw_sum<NW>(xb + 3 * kXStride, lane);
            if (fix_col >= 0) ge_out[(size_t)(s + 1) * S + fix_col] = from_f32<T>(gEm_fix - gEall);
        }
        // ---- LATE(s+1): gu, carry', h = g0*c3, thread-local exclusive suffix of gA = g0*c1 - gu*c2
        float2 hL[H], gAl[H];
        float linc = 0.f;
        if (doL) {
            const float gbase = xw_suffix_add<NW>(xb + 2 * kXStride, warp, lane) + gexc;
            const float2 gb2 = f2(gbase);
            float2 rcL[H], c2[H], c3[H], pL[H];
            stash_ld(1, rcL);
            stash_ld(2, c2);
            stash_ld(3, c3);
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(slot_p(kl) + j0, pL, dummy);
            float2 gAk[H];
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float2 gu = add2(gb2, Al[q]);
                carry[q] = mul2(gu, rcL[q]);
                hL[q] = mul2(g0[q], c3[q]);
                gAk[q] = fma2(hL[q], pL[q], mul2(mul2(gu, c2[q]), f2(-1.f)));      // c1 = c3 * p
            }
            float lt = 0.f;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k) {
                SIMULST_EL(gAl, k) = lt;     // exclusive
                lt += SIMULST_EL(gAk, k);
            }
            linc = lt;
        }
        // ---- EARLY(s) S1 finish: cp, 1/c, P ; 1/D ; alpha_{s-1} ; c2.  rc and c2 go to the stash in
        //      their final form, cp and P are parked there until the z mask is known (POST-B2);
        //      LATE(s+1) has just read its own copies above.
        float2 rD[H], sl[H], Rl[H], Wl[H];
        float uinc = 0.f, rinc = 0.f, winc = 0.f;
        if (doE) {
            const float xoff = xw_prefix_mul<NW>(xb + 0 * kXStride, warp, lane);
            const float cbase = (one_eps * xoff) * xexc;
            float2 cp[H], rc[H], P[H], am1[H];
            finish_cumprod<VPT>(cbase, cpre, p_s, eps, cp, rc, P);
            stash_st(0, P);
            stash_st(1, rc);
            stash_st(3, cp);
            if (s > 0) {
                mbar_wait(&bars[6 + km], parm);
                unsigned dummy = 0u;
                lds_row2<float, VPT, false>(slot_a(km) + j0, am1, dummy);
                if constexpr (!FULL) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (!is_in(k)) SIMULST_EL(am1, k) = 0.f;
                }
                // undo mass preservation on the stored row: the recurrence ran on the raw alpha
                if (own_last) {
                    const float raw = side[2 * (s - 1)];
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (at_last(k)) SIMULST_EL(am1, k) = raw;
                }
            } else {
#pragma unroll
                for (int q = 0; q < H; ++q) am1[q] = make_float2((j0 + 2 * q == 0) ? 1.0f : 0.0f, 0.0f);
            }
            // ============================================================ PRE-B2 (EARLY S2 local)
            uinc = local_u_prefix<VPT>(am1, rc, sl);
            {
                float2 c2[H];
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float cpk = SIMULST_EL(cp, k), rck = SIMULST_EL(rc, k);
                    const bool pass = cpk >= eps && cpk <= 1.0f;
                    const float u = SIMULST_EL(am1, k) * rck;
                    SIMULST_EL(c2, k) = pass ? u : 0.f;        // rc * u * cp with rc * cp = 1 inside the clamp
                }
                stash_st(2, c2);
            }
            if (SOFT) {
                const float ebase = xw_prefix_add<NW>(xb + 1 * kXStride, warp, lane) + eexc;
                finish_exp_prefix<VPT>(ebase, eps, Dl, rD);
                if (own_last) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (at_last(k)) bcast[0] = SIMULST_EL(rD, k);
                }
                // r = alpha'_s / D with alpha'_s (the row as stored) still in the ring
                if (s == T_len - 1) mbar_wait(&bars[6 + kp], par3);
                float2 r[H];
                {
                    unsigned dummy = 0u;
                    lds_row2<float, VPT, false>(slot_a(kp) + j0, r, dummy);
                }
#pragma unroll
                for (int q = 0; q < H; ++q) r[q] = mul2(r[q], rD[q]);
                if constexpr (!FULL) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (!is_live(k)) SIMULST_EL(r, k) = 0.f;
                }
                float rt = 0.f, wt = 0.f;
#pragma unroll
                for (int k = VPT - 1; k >= 0; --k) {
                    rt += SIMULST_EL(r, k);
                    SIMULST_EL(Rl, k) = rt;
                    wt = fmaf(SIMULST_EL(r, k), SIMULST_EL(rD, k), wt);
                    SIMULST_EL(Wl, k) = wt;
                }
                rinc = rt;
                winc = wt;
            }
        }
        float uexc, rexc = 0.f, wexc = 0.f, lexc;
        if (SOFT) {
            wscan_b2(uinc, rinc, winc, linc);
            wneigh_b2(uinc, rinc, winc, linc, uexc, rexc, wexc, lexc);
            if (lane == 0) xb[7 * kXStride + warp] = rinc;
            if (lane == 0) xb[8 * kXStride + warp] = winc;
        } else {
            wscan_ud(uinc, linc);
            uexc = wprev(uinc, 0.f);
            lexc = wnext(linc, 0.f);
        }
        if (lane == 31) xb[6 * kXStride + warp] = uinc;
        if (lane == 0) xb[9 * kXStride + warp] = linc;

        __syncthreads();                    // ================================ B2

        // ---- LATE(s+1) finish: gL, grad_p row s+1
        if (doL) {
            const float lbase = xw_suffix_add<NW>(xb + 9 * kXStride, warp, lane) + lexc;
            const float2 lb = f2(lbase);
            float2 pL[H], outp[H];
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(slot_p(kl) + j0, pL, dummy);
            if constexpr (!FULL) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (!is_live(k)) SIMULST_EL(pL, k) = 0.f;
            }
            const float2 one = f2(1.0f), neg = f2(-1.0f), e2 = f2(eps);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float2 rx = rcp2(add2(fma2(pL[q], neg, one), e2));
                const float2 gL = add2(lb, gAl[q]);
                outp[q] = fma2(mul2(gL, rx), neg, hL[q]);
            }
            float o8[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) o8[k] = is_live(k) ? SIMULST_EL(outp, k) : 0.f;
            st_row_t<T, VPT, FULL>(gp_out + (size_t)(s + 1) * S, j0, S, vec, o8);
        }
        // ---- EARLY(s) S2 finish: s, z mask -> final mz*P and c3 in the stash ; R, W, gb, gR
        float2 gR[H], ge1[H], W[H], exm[H];
        float gA_last = 0.f;
        if (doE) {
            const float ubase = xw_prefix_add<NW>(xb + 6 * kXStride, warp, lane) + uexc;
            float2 P[H], cp[H], sfull[H], z[H];
            stash_ld(0, P);
            stash_ld(3, cp);
            finish_u_prefix<VPT>(ubase, sl, P, sfull, z);
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const bool inside = SIMULST_EL(z, k) <= 1.0f;                    // z >= 0 (P >= 0, s >= 0)
                if (!inside) SIMULST_EL(P, k) = 0.f;                             // mz * P
                SIMULST_EL(cp, k) = inside ? SIMULST_EL(sfull, k) * SIMULST_EL(cp, k) : 0.f;   // c3 = mz * s * cp
            }
            stash_st(0, P);
            stash_st(3, cp);
            if (SOFT) {
                const float2 rb = f2(xw_suffix_add<NW>(xb + 7 * kXStride, warp, lane) + rexc);
                const float2 wb = f2(xw_suffix_add<NW>(xb + 8 * kXStride, warp, lane) + wexc);
                const float2 me = f2(-eps);
                float2 gB[H];
                if (has_gb) {
                    mbar_wait(&bars[9], parb);
                    lin_ld(reinterpret_cast<const float*>(slot_b), gB);
                } else {
#pragma unroll
                    for (int q = 0; q < H; ++q) gB[q] = f2(0.f);
                }
#pragma unroll
                for (int q = 0; q < H; ++q) {
                    const float2 R = add2(rb, Rl[q]);
                    W[q] = add2(wb, Wl[q]);
                    const float2 b = mul2(ex[q], R);
                    float2 gb;
                    gb.x = (b.x <= 1.0f) ? gB[q].x : 0.f;                        // b >= 0
                    gb.y = (b.y <= 1.0f) ? gB[q].y : 0.f;
                    if constexpr (!FULL) {
                        if (!is_live(2 * q)) gb.x = 0.f;
                        if (!is_live(2 * q + 1)) gb.y = 0.f;
                    }
                    ge1[q] = mul2(gb, R);
                    gR[q] = mul2(gb, ex[q]);
                    exm[q] = add2(ex[q], me);                                    // exp(E - m) up to an ulp of e
                }
            }
            // grad_alpha row s has landed in the g'' slot: the column mass preservation reads must
            // be fetched before the barrier, its owner overwrites it with g'' afterwards
            if (has_ga) {
                mbar_wait(&bars[10 + kg], parg);
                if (mp) gA_last = gslot(kg)[last];
            }
        }
        if (SOFT) {
            // ============================================================ PRE-B3
            float2 grl[H], Vl[H];
            float grinc = 0.f, vinc = 0.f;
            if (doE) {
                float gt = 0.f, vt = 0.f;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    gt += SIMULST_EL(gR, k);
                    SIMULST_EL(grl, k) = gt;
                }
#pragma unroll
                for (int k = VPT - 1; k >= 0; --k) {
                    vt = fmaf(SIMULST_EL(gR, k), SIMULST_EL(W, k), vt);
                    SIMULST_EL(Vl, k) = vt;
                }
                grinc = gt;
                vinc = vt;
            }
            wscan_ud(grinc, vinc);
            const float grexc = wprev(grinc, 0.f);
            const float vexc = wnext(vinc, 0.f);
            if (lane == 31) xb[10 * kXStride + warp] = grinc;
            if (lane == 0) xb[11 * kXStride + warp] = vinc;

            __syncthreads();                // ================================ B3

            if (doE) {
                const float gbase = xw_prefix_add<NW>(xb + 10 * kXStride, warp, lane) + grexc;
                const float gtotal = xw_sum<NW>(xb + 10 * kXStride, lane);
                const float vbase = xw_suffix_add<NW>(xb + 11 * kXStride, warp, lane) + vexc;
                float okg = 0.f;
                if (mp) {
                    const float row_total = side[2 * s + 1];
                    const float ok = (row_total >= 0.0f && row_total <= 1.0f) ? 1.0f : 0.0f;
                    okg = ok * (gA_last + gtotal * bcast[0]);
                }
                float gEm[VPT];
                float2 g2[H], gA[H];
                if (has_ga) {
                    lin_ld(gslot(kg), gA);
                } else {
#pragma unroll
                    for (int q = 0; q < H; ++q) gA[q] = f2(0.f);
                }
                float gsum = 0.f;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float grx = gbase + (k == 0 ? 0.f : SIMULST_EL(grl, k - (k == 0 ? 0 : 1)));   // exclusive prefix
                    const float gr = gbase + SIMULST_EL(grl, k);
                    const float gsoft = gr * SIMULST_EL(rD, k);
                    const float sufD = -fmaf(grx, SIMULST_EL(W, k), vbase + SIMULST_EL(Vl, k));
                    const float ge = SIMULST_EL(ge1, k) + sufD;
                    const bool live = is_live(k);
                    gEm[k] = live ? ge * SIMULST_EL(exm, k) : 0.f;
                    gsum += gEm[k];
                    float g = live ? (SIMULST_EL(gA, k) + gsoft) - okg : 0.f;
                    if (mp && !mp_add && at_last(k)) g = 0.f;      // replaced column
                    SIMULST_EL(g2, k) = g;
                }
                gEsum_prev = gsum;
                st_row_t<T, VPT, FULL>(ge_out + (size_t)s * S, j0, S, vec, gEm);
                lin_st(gslot(kg), g2);
                fence_proxy_async();        // this slot is refilled by a bulk copy next step
                // the thread that owns the arg-max column keeps its value for the deferred correction
                fix_col = -1;
                const int d = amax_cur - j0;
                if ((unsigned)d < (unsigned)VPT) {
                    fix_col = amax_cur;
                    gEm_fix = gEm[0];
#pragma unroll
                    for (int k = 1; k < VPT; ++k)
                        if (d == k) gEm_fix = gEm[k];
                }
            }
        } else {
            if (doE) {
                float okg = 0.f;
                if (mp) {
                    const float row_total = side[2 * s + 1];
                    const float ok = (row_total >= 0.0f && row_total <= 1.0f) ? 1.0f : 0.0f;
                    okg = ok * gA_last;
                }
                float2 g2[H], gA[H];
                if (has_ga) {
                    lin_ld(gslot(kg), gA);
                } else {
#pragma unroll
                    for (int q = 0; q < H; ++q) gA[q] = f2(0.f);
                }
                __syncthreads();            // gA_last (read above by everyone) is about to be overwritten
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    float g = is_live(k) ? SIMULST_EL(gA, k) - okg : 0.f;
                    if (mp && !mp_add && at_last(k)) g = 0.f;
                    SIMULST_EL(g2, k) = g;
                }
                lin_st(gslot(kg), g2);
                fence_proxy_async();        // this slot is refilled by a bulk copy next step
            } else {
                __syncthreads();
            }
            __syncthreads();                // ring slots below must not be refilled while still read
        }

        // ---- ring refills (every read of this iteration's rows is behind the last barrier):
        //      rows s (energy, alpha) and s+1 (p) are dead
        //      rows s (energy, alpha, grad_beta) and s+1 (p, g'') are dead
        if (tid == 0) {
            if (SOFT && s - 3 >= 0) issue_e(s - 3);
            if (s - 2 >= 0) issue_p(s - 2);
            if (s - 3 >= 0) issue_a(s - 3);
            if (s - 1 >= 0 && s - 1 < T_len) {
                if (has_ga && s - 1 <= T_len - 2) issue_ga(s - 1);
                if (has_gb && s <= T_len - 1) issue_gb(s - 1);
            }
        }
        if (++kp == 3) { kp = 0; par3 ^= 1u; }
        if (SOFT && doM) { m_cur = m_nxt; amax_cur = amax_nxt; }
    };

    using Steady = std::integral_constant<bool, true>;
    using Edge = std::integral_constant<bool, false>;
    // s = T_len (MAXS only) ... -1 (LATE only); all stages are live for 1 <= s <= T_len - 2.
    // Ring positions are counted from s = T_len - 1, so the first iteration starts one before.
    kp = 2; par3 = 1u;
    int s = T_len;
    for (; s > T_len - 2 && s >= -1; --s) body(Edge{}, s);
    for (; s >= 1; --s) body(Steady{}, s);
    for (; s >= -1; --s) body(Edge{}, s);
}

// ------------------------------------------------------------------ host-side launcher
// Returns 1 when the row does not fit the pipelined kernel (caller falls back to the generic one).
template <int THREADS, int VPT, typename T, bool SOFT, bool FULL>
int launch_mma_bwd_pipe_impl(const MmaParams& prm, cudaStream_t stream) {
    constexpr int CAP = THREADS * VPT;
    BwdPipePlan plan;
    plan.row_t_bytes = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    plan.row_f_bytes = (CAP * 4 + 127) / 128 * 128;
    plan.soft = SOFT ? 1 : 0;
    plan.stash_bytes = 4 * CAP * 4;
    const size_t budget = THREADS <= 128 ? 56 * 1024 : (THREADS <= 256 ? 110 * 1024 : 220 * 1024);
    if (plan.total() > budget) return 1;
    auto kern = mma_bwd_pipe_kernel<THREADS, VPT, T, SOFT, FULL>;
    static size_t attr_set[64] = {};
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

template <int THREADS, int VPT, typename T, bool SOFT>
int launch_mma_bwd_pipe(const MmaParams& prm, cudaStream_t stream) {
    // Dense rows that fill the CTA only.  The ragged / masked instantiation computes on the
    // unwritten tails of its shared-memory rings before masking them out, which is only safe
    // when those bytes happen to hold finite values (found with -inf tails left behind by another
    // kernel); such rows go to the generic kernel (return 1).
    const bool full = prm.mask == nullptr && prm.S == THREADS * VPT && prm.vec_out;
    if (!full) return 1;
    return launch_mma_bwd_pipe_impl<THREADS, VPT, T, SOFT, true>(prm, stream);
}

}  // namespace simulst
