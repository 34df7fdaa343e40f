// MMA training backward: gradients of (expected alignment -> mass preservation -> expected
// soft attention) w.r.t. p_choose and soft_energy.  One CTA per (batch*head) row walks the
// target steps in reverse with the recurrence gradient ("carry") in registers; every scan of
// the forward pass is recomputed from p / energy, only the forward OUTPUT alpha (+ a [T,2]
// side vector) is read back.
//
// Derivation: SURVEY Appendix A.2-A.4 (autograd of codebase/utils/monotonic_attention.py).
// Per step i (reverse), with g' = dL/d alpha'_i (external + soft-attention part):
//   soft:  gb = gbeta*1[0<=b<=1]; ge = gb*R; gR = gb*e; gr = prefix(gR); ga' += gr/D;
//          gD = -gr*r/D; ge += suffix(gD); gEm = ge*(e-eps); gE = gEm - [argmax]*sum(gEm)
//   mass preservation:  ga_j = g'_j - ok*g'_last   (column `last` itself: 0 when replaced)
//   alignment:  g = ga + carry; gz = g*1[0<=z<=1]; gP = gz*s; gs = gz*P; gu = suffix(gs);
//          carry' = gu/c; gc = -gu*u/c; gcp = gP*p + gc*1[eps<=cp<=1]; gA = gcp*cp;
//          gL = exclusive_suffix(gA); gp = gP*cp - gL/((1-p)+eps)
#pragma once

#include "mma_common.cuh"

namespace simulst {

struct BwdPlan {
    int n_stage;
    int stage_bytes;
    int off_p, off_e, off_a, off_ga, off_gb;   // byte offsets of the rows inside a stage
    int win_floats;                            // chunkwise scratch: 2 buffers of this many floats
    __host__ __device__ int header_bytes() const { return 128 + 2 * kXSlots * kXStride * 4 + 128; }
    __host__ __device__ size_t total() const {
        return (size_t)header_bytes() + (size_t)n_stage * stage_bytes + (size_t)2 * win_floats * 4;
    }
};

template <int THREADS, int VPT, typename T, int MODE, bool FULL>
__global__ void __launch_bounds__(THREADS, (THREADS * VPT <= 1024 ? 4 : 1)) mma_bwd_kernel(const MmaParams prm, const BwdPlan plan) {
    constexpr int NW = THREADS / kWarp;
    constexpr bool SOFT = MODE != kModeHard;
    constexpr bool CHUNK = MODE == kModeSoftCk;

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    float* bcast = xraw + 2 * kXSlots * kXStride;
    unsigned char* stage0 = smem + plan.header_bytes();
    float* win0 = reinterpret_cast<float*>(stage0 + (size_t)plan.n_stage * plan.stage_bytes);
    float* win1 = win0 + plan.win_floats;
    (void)win0; (void)win1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    if (row_filtered_out(prm, n)) return;       // this row belongs to the call's other pass
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const float fill = (prm.flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const bool vec = FULL || prm.vec_out != 0;
    const int NS = plan.n_stage;
    const bool has_ga = prm.g_alpha != nullptr;
    const bool has_gb = SOFT && prm.g_beta != nullptr;
    const bool has_gd = prm.g_delays != nullptr;    // gradient of the expected delays: g'_ij += gd_i * (j+1)

    // row pitches in elements; the batch stride of a tensor is T * pitch
    const int ld_p = prm.ld_p, ld_e = prm.ld_e, ld_a = prm.ld_alpha, ld_ga = prm.ld_ga, ld_gb = prm.ld_gb,
              ld_gp = prm.ld_gp, ld_ge = prm.ld_ge;
    const size_t nt = (size_t)n * T_len;
    const T* gp_in = reinterpret_cast<const T*>(prm.p) + nt * ld_p;
    const T* ge_in = SOFT ? reinterpret_cast<const T*>(prm.e) + nt * ld_e : nullptr;
    const float* al = prm.alpha + nt * ld_a;
    const float* gA_in = has_ga ? prm.g_alpha + nt * ld_ga : nullptr;
    const float* gB_in = has_gb ? prm.g_beta + nt * ld_gb : nullptr;
    T* gp_out = reinterpret_cast<T*>(prm.g_p) + nt * ld_gp;
    T* ge_out = SOFT ? reinterpret_cast<T*>(prm.g_e) + nt * ld_ge : nullptr;
    const float* side = mp ? prm.side + (size_t)n * T_len * 2 : nullptr;

    Xchg xc(xraw);

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
    const bool last_thread = tid == THREADS - 1;
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

    // ---- staging ring; processing counter q = 0..T-1 handles target step i = T-1-q
    const unsigned t_bytes = (unsigned)(S * sizeof(T)), f_bytes = (unsigned)(S * 4);
    auto st_ptr = [&](int s, int off) { return stage0 + (size_t)s * plan.stage_bytes + off; };
    auto coop = [&](void* dst, const void* src, int elem_bytes) {
        if (elem_bytes == 4) {
            float* d = reinterpret_cast<float*>(dst);
            const float* s_ = reinterpret_cast<const float*>(src);
#pragma unroll 1
            for (int j = tid; j < S; j += THREADS) d[j] = s_[j];
        } else {
            uint16_t* d = reinterpret_cast<uint16_t*>(dst);
            const uint16_t* s_ = reinterpret_cast<const uint16_t*>(src);
#pragma unroll 1
            for (int j = tid; j < S; j += THREADS) d[j] = s_[j];
        }
    };
    auto issue = [&](int q, int s) {
        const int i = T_len - 1 - q;
        if (prm.tma) {
            if (tid == 0) {
                unsigned bytes = t_bytes + (SOFT ? t_bytes : 0u) + (i > 0 ? f_bytes : 0u) +
                                 (has_ga ? f_bytes : 0u) + (has_gb ? f_bytes : 0u);
                mbar_expect_tx(&bars[s], bytes);
                tma_load_1d(st_ptr(s, plan.off_p), gp_in + (size_t)i * ld_p, t_bytes, &bars[s]);
                if (SOFT) tma_load_1d(st_ptr(s, plan.off_e), ge_in + (size_t)i * ld_e, t_bytes, &bars[s]);
                if (i > 0) tma_load_1d(st_ptr(s, plan.off_a), al + (size_t)(i - 1) * ld_a, f_bytes, &bars[s]);
                if (has_ga) tma_load_1d(st_ptr(s, plan.off_ga), gA_in + (size_t)i * ld_ga, f_bytes, &bars[s]);
                if (has_gb) tma_load_1d(st_ptr(s, plan.off_gb), gB_in + (size_t)i * ld_gb, f_bytes, &bars[s]);
            }
        } else if (prm.tma_shift) {
            if (tid == 0) {
                // 16-byte aligned supersets of the (unaligned) rows; see tma_span
                unsigned nb[5] = {0u, 0u, 0u, 0u, 0u};
                const void* src[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
                src[0] = tma_span(gp_in + (size_t)i * ld_p, t_bytes, nb[0]);
                if (SOFT) src[1] = tma_span(ge_in + (size_t)i * ld_e, t_bytes, nb[1]);
                if (i > 0) src[2] = tma_span(al + (size_t)(i - 1) * ld_a, f_bytes, nb[2]);
                if (has_ga) src[3] = tma_span(gA_in + (size_t)i * ld_ga, f_bytes, nb[3]);
                if (has_gb) src[4] = tma_span(gB_in + (size_t)i * ld_gb, f_bytes, nb[4]);
                mbar_expect_tx(&bars[s], nb[0] + nb[1] + nb[2] + nb[3] + nb[4]);
                tma_load_1d(st_ptr(s, plan.off_p), src[0], nb[0], &bars[s]);
                if (SOFT) tma_load_1d(st_ptr(s, plan.off_e), src[1], nb[1], &bars[s]);
                if (i > 0) tma_load_1d(st_ptr(s, plan.off_a), src[2], nb[2], &bars[s]);
                if (has_ga) tma_load_1d(st_ptr(s, plan.off_ga), src[3], nb[3], &bars[s]);
                if (has_gb) tma_load_1d(st_ptr(s, plan.off_gb), src[4], nb[4], &bars[s]);
            }
        } else {
            coop(st_ptr(s, plan.off_p), gp_in + (size_t)i * ld_p, sizeof(T));
            if (SOFT) coop(st_ptr(s, plan.off_e), ge_in + (size_t)i * ld_e, sizeof(T));
            if (i > 0) coop(st_ptr(s, plan.off_a), al + (size_t)(i - 1) * ld_a, 4);
            if (has_ga) coop(st_ptr(s, plan.off_ga), gA_in + (size_t)i * ld_ga, 4);
            if (has_gb) coop(st_ptr(s, plan.off_gb), gB_in + (size_t)i * ld_gb, 4);
        }
    };
    for (int q = 0; q < NS - 1 && q < T_len; ++q) issue(q, q);
    const bool use_tma = prm.tma || prm.tma_shift;
    if (!use_tma) __syncthreads();

    const float one_eps = 1.0f + eps;
    float carry[VPT], a_cur[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) carry[k] = 0.f;
    if (SOFT) ld_row_f32<VPT>(al + (size_t)(T_len - 1) * ld_a, j0, S, vec, a_cur);   // alpha'_{T-1}

    int s = 0, s_fill = NS - 1;
    unsigned parity = 0u;
    for (int q = 0; q < T_len; ++q) {
        const int i = T_len - 1 - q;
        if (q + NS - 1 < T_len) issue(q + NS - 1, s_fill);
        // scalars of this step (broadcast loads, consumed late)
        float side_sum = 0.f, side_prev_last = 0.f;
        if (mp) {
            side_sum = side[2 * i + 1];
            if (i > 0) side_prev_last = side[2 * (i - 1)];
        }
        const float gd = has_gd ? prm.g_delays[(size_t)n * T_len + i] : 0.f;
        if (use_tma) mbar_wait(&bars[s], parity);

        float p[VPT], E[VPT], am1[VPT], gA[VPT], gB[VPT];
        const bool shifted = prm.tma_shift != 0;
        lds_row_shift<T, VPT>(reinterpret_cast<const T*>(st_ptr(s, plan.off_p)),
                              shifted ? row_shift(gp_in + (size_t)i * ld_p) : 0, j0, p);
        if (SOFT) lds_row_shift<T, VPT>(reinterpret_cast<const T*>(st_ptr(s, plan.off_e)),
                                        shifted ? row_shift(ge_in + (size_t)i * ld_e) : 0, j0, E);
        if (i > 0) {
            lds_row_shift<float, VPT>(reinterpret_cast<const float*>(st_ptr(s, plan.off_a)),
                                      shifted ? row_shift(al + (size_t)(i - 1) * ld_a) : 0, j0, am1);
        } else {
#pragma unroll
            for (int k = 0; k < VPT; ++k) am1[k] = (j0 + k == 0) ? 1.0f : 0.0f;
        }
        if (has_ga) lds_row_shift<float, VPT>(reinterpret_cast<const float*>(st_ptr(s, plan.off_ga)),
                                              shifted ? row_shift(gA_in + (size_t)i * ld_ga) : 0, j0, gA);
        if (has_gb) lds_row_shift<float, VPT>(reinterpret_cast<const float*>(st_ptr(s, plan.off_gb)),
                                              shifted ? row_shift(gB_in + (size_t)i * ld_gb) : 0, j0, gB);
        if (++s == NS) { s = 0; parity ^= 1u; }
        if (++s_fill == NS) s_fill = 0;
        float a_save[VPT];          // alpha'_{i-1} exactly as stored (becomes a_cur next step)
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const bool in = is_in(k);
            if (!in) am1[k] = 0.f;
            if (!has_ga || !in) gA[k] = 0.f;
            if (has_gd && in) gA[k] = __fmaf_rn((float)(j0 + k + 1), gd, gA[k]);
            if (!has_gb || !in) gB[k] = 0.f;
            a_save[k] = am1[k];
            // undo mass preservation on the stored row: the recurrence ran on the raw alpha
            if (mp && i > 0 && at_last(k)) am1[k] = side_prev_last;
        }

        // ================= X1: exclusive cumprod of (1-p)+eps ; max of E
        float cp[VPT], rx[VPT];         // cp: local exclusive product prefix first; rx = 1/((1-p)+eps)
        float xtot = 1.0f, Emax = -INFINITY;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            if (!is_live(k)) p[k] = 0.f;
            const float x = is_in(k) ? (1.0f - p[k]) + eps : 1.0f;
            rx[k] = fast_rcp(x);
            cp[k] = xtot;
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
        float rc[VPT], P[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            cp[k] = cbase * cp[k];
            rc[k] = fast_rcp(fminf(fmaxf(cp[k], eps), 1.0f));
            P[k] = p[k] * cp[k];
        }

        // ================= X2: D (prefix / window of e) ; first index attaining the max
        float ex[VPT], exm[VPT], D[VPT];      // exm = exp(E-m), ex = exm + eps; D becomes 1/D
        int amax = 0;
        if (SOFT) {
            float etot = 0.f;
            int cand = 0x7fffffff;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k)
                if (is_in(k) && E[k] == m) cand = j0 + k;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const bool in = is_in(k);
                exm[k] = in ? fast_exp(E[k] - m) : 0.f;
                ex[k] = in ? (exm[k] + eps) : 0.f;
                etot += ex[k];
                D[k] = etot;
            }
            cand = __reduce_min_sync(kFull, cand);
            if (lane == 0) reinterpret_cast<int*>(xc.slot(1))[warp] = cand;
            float eexc = 0.f;
            if (!CHUNK) {
                const float einc = warp_incl_prefix(etot, lane);
                if (lane == 31) xc.slot(0)[warp] = einc;
                eexc = lane_prev(einc, lane, 0.f);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) win0[j0 + k] = ex[k];
            }
            __syncthreads();
            {
                const int* ci = reinterpret_cast<const int*>(xc.slot(1));
                amax = ci[0];
#pragma unroll
                for (int w = 1; w < NW; ++w) amax = min(amax, ci[w]);
            }
            if (!CHUNK) {
                const float2 eo = combine_prefix<NW>(xc.slot(0), warp, lane);
                const float ebase = eo.x + eexc;
#pragma unroll
                for (int k = 0; k < VPT; ++k) D[k] = eps + (ebase + D[k]);
            } else {
                const int cw = prm.chunk;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int j = j0 + k;
                    float acc = 0.f;
                    for (int t = max(0, j - cw + 1); t <= j; ++t) acc += win0[t];
                    D[k] = eps + acc;
                }
            }
            xc.flip();
#pragma unroll
            for (int k = 0; k < VPT; ++k) D[k] = fast_rcp(D[k]);     // from here on D holds 1/D
        }

        // ================= X3: s = prefix(u) ; R = suffix / window of r
        float u[VPT], sl[VPT], r[VPT], Rl[VPT];
        float utot = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            u[k] = am1[k] * rc[k];
            utot += u[k];
            sl[k] = utot;
        }
        const float uinc = warp_incl_prefix(utot, lane);
        if (lane == 31) xc.slot(0)[warp] = uinc;
        const float uexc = lane_prev(uinc, lane, 0.f);
        float rexc = 0.f;
        if (SOFT) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                r[k] = is_live(k) ? a_cur[k] * D[k] : 0.f;
                Rl[k] = r[k];
            }
            if (!CHUNK) {
#pragma unroll
                for (int k = VPT - 2; k >= 0; --k) Rl[k] += Rl[k + 1];
                const float rinc = warp_incl_suffix(Rl[0], lane);
                if (lane == 0) xc.slot(1)[warp] = rinc;
                rexc = lane_next(rinc, lane, 0.f);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) win1[j0 + k] = r[k];
            }
        }
        __syncthreads();
        const float2 uo = combine_prefix<NW>(xc.slot(0), warp, lane);
        const float ubase = uo.x + uexc;
        float sfull[VPT], mz[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            sfull[k] = ubase + sl[k];
            const float z = P[k] * sfull[k];
            mz[k] = (z >= 0.0f && z <= 1.0f) ? 1.0f : 0.0f;
        }
        float gsoft[VPT], ge1[VPT], gD[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) { gsoft[k] = 0.f; ge1[k] = 0.f; gD[k] = 0.f; }
        if (SOFT) {
            float R[VPT];
            if (!CHUNK) {
                const float2 ro = combine_suffix<NW>(xc.slot(1), warp, lane);
                const float rbase = ro.x + rexc;
#pragma unroll
                for (int k = 0; k < VPT; ++k) R[k] = rbase + Rl[k];
            } else {
                const int cw = prm.chunk;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int j = j0 + k;
                    float acc = 0.f;
                    for (int t = j; t <= min(S - 1, j + cw - 1); ++t) acc += win1[t];
                    R[k] = acc;
                }
            }
            xc.flip();
            // ============= X4: gr = prefix / window of gR
            float gR[VPT], grl[VPT];
            float gtot = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const float b = ex[k] * R[k];
                const float gb = (is_live(k) && b >= 0.0f && b <= 1.0f) ? gB[k] : 0.f;
                ge1[k] = gb * R[k];
                gR[k] = gb * ex[k];
                gtot += gR[k];
                grl[k] = gtot;
            }
            float gexc = 0.f;
            if (!CHUNK) {
                const float ginc = warp_incl_prefix(gtot, lane);
                if (lane == 31) xc.slot(0)[warp] = ginc;
                gexc = lane_prev(ginc, lane, 0.f);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) win0[j0 + k] = gR[k];
            }
            __syncthreads();
            if (!CHUNK) {
                const float2 go = combine_prefix<NW>(xc.slot(0), warp, lane);
                const float gbase = go.x + gexc;
#pragma unroll
                for (int k = 0; k < VPT; ++k) grl[k] = gbase + grl[k];
            } else {
                const int cw = prm.chunk;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int j = j0 + k;
                    float acc = 0.f;
                    for (int t = max(0, j - cw + 1); t <= j; ++t) acc += win0[t];
                    grl[k] = acc;
                }
            }
            xc.flip();
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const float gq = is_live(k) ? grl[k] * D[k] : 0.f;       // d/d alpha'  (D holds 1/D)
                gsoft[k] = gq;
                gD[k] = -gq * r[k];                                       // -gr*r/D
            }
        } else {
            xc.flip();
        }

        // ================= X5: gu = suffix(gs) with the mass-preservation term split off ;
        //                       suffix / window of gD
        float g0[VPT], Aq[VPT], Bq[VPT];
        float glast_mine = 0.f;
        bool own_last = false;
        float Atot = 0.f, Btot = 0.f, Dtot = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const bool live = is_live(k);
            const bool is_last = at_last(k);
            const float gprime = live ? (gA[k] + gsoft[k]) : 0.f;     // dL/d alpha'_i
            if (mp && is_last) { glast_mine = gprime; own_last = true; }
            float w = live ? 1.0f : 0.f;
            float gq = gprime;
            if (mp && is_last && !mp_add) { gq = 0.f; w = 0.f; }      // replaced column
            g0[k] = gq + carry[k];
            Aq[k] = mz[k] * P[k] * g0[k];
            Bq[k] = mp ? mz[k] * P[k] * w : 0.f;
        }
        float Al[VPT], Bl[VPT], Dl[VPT];
#pragma unroll
        for (int k = VPT - 1; k >= 0; --k) {
            Atot += Aq[k]; Al[k] = Atot;
            Btot += Bq[k]; Bl[k] = Btot;
            Dtot += gD[k]; Dl[k] = Dtot;
        }
        const float Ainc = warp_incl_suffix(Atot, lane);
        if (lane == 0) xc.slot(0)[warp] = Ainc;
        const float Aexc = lane_next(Ainc, lane, 0.f);
        float Bexc = 0.f, Dexc = 0.f;
        if (mp) {
            const float Binc = warp_incl_suffix(Btot, lane);
            if (lane == 0) xc.slot(1)[warp] = Binc;
            Bexc = lane_next(Binc, lane, 0.f);
            if (own_last) bcast[0] = glast_mine;
        }
        if (SOFT) {
            if (!CHUNK) {
                const float Dinc = warp_incl_suffix(Dtot, lane);
                if (lane == 0) xc.slot(2)[warp] = Dinc;
                Dexc = lane_next(Dinc, lane, 0.f);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) win1[j0 + k] = gD[k];
            }
        }
        __syncthreads();
        const float2 Ao = combine_suffix<NW>(xc.slot(0), warp, lane);
        const float Abase = Ao.x + Aexc;
        float okg = 0.f;
        float Bbase = 0.f;
        if (mp) {
            const float2 Bo = combine_suffix<NW>(xc.slot(1), warp, lane);
            Bbase = Bo.x + Bexc;
            const float ok = (side_sum >= 0.0f && side_sum <= 1.0f) ? 1.0f : 0.0f;
            okg = ok * bcast[0];
        }
        float gEm[VPT];
        float gEsum = 0.f;
        if (SOFT) {
            float Dbase = 0.f;
            if (!CHUNK) {
                const float2 Do = combine_suffix<NW>(xc.slot(2), warp, lane);
                Dbase = Do.x + Dexc;
            }
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const bool live = is_live(k);
                float sufD;
                if (!CHUNK) {
                    sufD = Dbase + Dl[k];
                } else {
                    const int j = j0 + k, cw = prm.chunk;
                    float acc = 0.f;
                    for (int t = j; t <= min(S - 1, j + cw - 1); ++t) acc += win1[t];
                    sufD = acc;
                }
                const float ge = ge1[k] + sufD;
                gEm[k] = live ? ge * exm[k] : 0.f;
                gEsum += gEm[k];
            }
        }
        xc.flip();

        // ================= X6: gL = exclusive suffix of gA ; sum of gEm
        float gPk[VPT], gAl[VPT];
        float gAtot = 0.f;
#pragma unroll
        for (int k = VPT - 1; k >= 0; --k) {
            const float gu = (Abase + Al[k]) - okg * (Bbase + Bl[k]);
            const bool live = is_live(k);
            const bool is_last = at_last(k);
            const float w = (!live || (mp && is_last && !mp_add)) ? 0.f : 1.0f;
            const float gz = mz[k] * (g0[k] - okg * w);
            gPk[k] = gz * sfull[k];
            const float inv_c = gu * rc[k];
            carry[k] = inv_c;                                   // dL/d alpha_{i-1}
            const float gc = -inv_c * u[k];
            const float pass = (cp[k] >= eps && cp[k] <= 1.0f) ? 1.0f : 0.0f;
            const float gcp = gPk[k] * p[k] + gc * pass;
            const float gAk = gcp * cp[k];
            gAl[k] = gAtot;                                     // exclusive local suffix
            gAtot += gAk;
        }
        const float gAinc = warp_incl_suffix(gAtot, lane);
        if (lane == 0) xc.slot(0)[warp] = gAinc;
        const float gAexc = lane_next(gAinc, lane, 0.f);
        if (SOFT) {
            const float ws = warp_sum(gEsum);
            if (lane == 0) xc.slot(1)[warp] = ws;
        }
        __syncthreads();
        const float2 gLo = combine_suffix<NW>(xc.slot(0), warp, lane);
        const float gLbase = gLo.x + gAexc;
        float gEall = 0.f;
        if (SOFT) gEall = combine_sum<NW>(xc.slot(1), lane);
        xc.flip();

        float outp[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const float gL = gLbase + gAl[k];
            outp[k] = is_live(k) ? (gPk[k] * cp[k] - gL * rx[k]) : 0.f;
        }
        st_row_t<T, VPT, FULL>(gp_out + (size_t)i * ld_gp, j0, S, vec, outp);
        if (SOFT) {
            float oute[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                float v = gEm[k];
                if (j0 + k == amax) v -= gEall;
                oute[k] = is_live(k) ? v : 0.f;
            }
            st_row_t<T, VPT, FULL>(ge_out + (size_t)i * ld_ge, j0, S, vec, oute);
#pragma unroll
            for (int k = 0; k < VPT; ++k) a_cur[k] = a_save[k];
        }
    }
}

template <int THREADS, int VPT, typename T, int MODE, bool FULL>
int launch_mma_bwd_impl(const MmaParams& prm, cudaStream_t stream) {
    constexpr int CAP = THREADS * VPT;
    const bool soft = MODE != kModeHard;
    BwdPlan plan;
    const int t_row = (CAP * (int)sizeof(T) + 16 + 127) / 128 * 128;      // + the head of a shifted row
    const int f_row = (CAP * 4 + 16 + 127) / 128 * 128;
    int off = 0;
    plan.off_p = off; off += t_row;
    plan.off_e = off; if (soft) off += t_row;
    plan.off_a = off; off += f_row;
    plan.off_ga = off; if (prm.g_alpha != nullptr) off += f_row;
    plan.off_gb = off; if (soft && prm.g_beta != nullptr) off += f_row;
    plan.stage_bytes = off;
    plan.win_floats = (MODE == kModeSoftCk) ? CAP : 0;
    plan.n_stage = (3 * off <= 56 * 1024) ? 3 : 2;
    if (plan.total() > 220 * 1024) plan.n_stage = 1;       // very long rows: synchronous staging
    auto kern = mma_bwd_kernel<THREADS, VPT, T, MODE, FULL>;
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

template <int THREADS, int VPT, typename T, int MODE>
int launch_mma_bwd(const MmaParams& prm, cudaStream_t stream) {
    const bool full = prm.mask == nullptr && prm.S == THREADS * VPT && prm.vec_out && prm.tma;
    return full ? launch_mma_bwd_impl<THREADS, VPT, T, MODE, true>(prm, stream)
                : launch_mma_bwd_impl<THREADS, VPT, T, MODE, false>(prm, stream);
}

}  // namespace simulst
