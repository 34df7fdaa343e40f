// MMA training forward for LONG rows in SMALL batches: one thread-block CLUSTER per (batch*head) row.
//
// One CTA per row (mma_fwd_pipe.cuh) leaves most of the GPU dark when there are fewer rows than SMs
// (SURVEY's long-form configuration: 64 rows), and from ~2048 source frames on the one SM a row lives on is
// issue bound: it runs 8 - 16 warps through every step (2.0 us per step at S = 4096 against a 0.8 us latency
// floor).  Here CL = 2 / 4 / 8 CTAs of 96 or 128 threads share a row: CTA `rank` owns columns
// [rank * SLICE, (rank + 1) * SLICE), stages its own slice of every input row by TMA, and the per-warp scan
// totals that the single-CTA kernel exchanges through shared memory + __syncthreads() are exchanged
// CLUSTER-wide: every warp writes its values into the exchange buffer of all CL CTAs (st.async to the peers'
// distributed shared memory, completing transaction bytes on the peer's mbarrier), arrives on its own CTA's
// mbarrier, and waits there -- no barrier.cluster in the step loop (the probe in tests/probes/cluster_probe.cu
// prices this exchange at ~0.07 us against 0.27 us for a cluster barrier).  Same 4-stage software pipeline, same
// arithmetic in the same order as mma_fwd_pipe_kernel (cross-warp offsets are combined over CL * NW warps in
// cluster order), so the results are bit-identical to the single-CTA kernel's.
//
// Scope: hard / infinite-lookback attention, no padding mask, dense 16-byte rows, S a multiple of VPT.
#pragma once

#include "mma_fwd_pipe.cuh"

namespace simulst {

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(remote_bar) : "memory");
}
// Cross-warp combines over an exchange buffer laid out [warp][8 values] (a warp's eight values travel to a
// peer CTA as two 16-byte stores): value `slot` of warp w sits at wt[w * 8].  Same association as the
// xw_* helpers of mma_scan.cuh (sequential in warp order up to 8 warps, shuffle scans beyond).
constexpr int kXW = 8;
template <int NW>
__device__ __forceinline__ float cw_prefix_add(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW <= 8) {
        float off = 0.f;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) {
            const float v = wt[w * kXW];
            if (w < warp) off += v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane * kXW] : 0.f;
        const float inc = wscan_prefix_add(v);
        return __shfl_sync(kFull, wprev(inc, 0.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float cw_prefix_mul(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW <= 8) {
        float off = 1.f;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) {
            const float v = wt[w * kXW];
            if (w < warp) off *= v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane * kXW] : 1.f;
        const float inc = wscan_prefix_mul(v);
        return __shfl_sync(kFull, wprev(inc, 1.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float cw_suffix_add(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW <= 8) {
        float off = 0.f;
#pragma unroll
        for (int w = NW - 1; w > 0; --w) {
            const float v = wt[w * kXW];
            if (w > warp) off += v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane * kXW] : 0.f;
        const float inc = wscan_suffix_add(v);
        return __shfl_sync(kFull, wnext(inc, 0.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float cw_sum(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float tot = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) tot += wt[w * kXW];
        return tot;
    } else {
        return warp_sum((lane < NW) ? wt[lane * kXW] : 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float cw_max(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float m = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, wt[w * kXW]);
        return m;
    } else {
        return wmax_redux((lane < NW) ? wt[lane * kXW] : -INFINITY);
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_release(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=: mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

template <int THREADS, int VPT, typename T, bool SOFT, int CL>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 4 : 2))
mma_fwd_cluster_kernel(const MmaParams prm) {
    using PS = PipeStatic<THREADS, VPT, T, SOFT>;
    constexpr int NW = THREADS / kWarp;
    constexpr int NWC = NW * CL;                            // warps of the whole row, in cluster order
    constexpr int kIssuers = (SOFT && NW > 1) ? 2 : 1;
    constexpr int H = VPT / 2;
    constexpr int SLICE = THREADS * VPT;
    constexpr int NS = PS::kNS;
    static_assert(NWC <= kXStride && VPT % 4 == 0 && NS >= 2, "cluster forward: at most 32 warps per row");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // [0, NS): ring; [4], [5]: exchange (two phases in flight)
    uint64_t* xbar = bars + 4;
    float* xbuf = reinterpret_cast<float*>(smem + 128);               // [2][kPipeSlots][32]
    unsigned char* stage0 = smem + PS::kHeader;
    float4* stash = reinterpret_cast<float4*>(stage0 + NS * PS::kRows * PS::kRowBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(kFull, tid >> 5, 0);
    const unsigned rank = cluster_ctarank();
    const int n = blockIdx.x / CL;
    const int gw = (int)rank * NW + warp;                               // this warp's index in the row
    const int S = prm.S, T_len = prm.T;
    const int j0 = ((int)rank * THREADS + tid) * VPT;                   // first column of this thread
    const int l0 = tid * VPT;                                           // ... inside the CTA's slice
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const int slice_base = (int)rank * SLICE;
    const int slice_n = max(0, min(SLICE, S - slice_base));             // live columns of this CTA's slice
    const bool has_data = slice_n > 0;

    const T* gp = reinterpret_cast<const T*>(prm.p) + (size_t)n * T_len * S;
    const T* ge = SOFT ? reinterpret_cast<const T*>(prm.e) + (size_t)n * T_len * S : nullptr;
    float* g_alpha = prm.alpha + (size_t)n * T_len * S;
    float* g_beta = SOFT ? prm.beta + (size_t)n * T_len * S : nullptr;

    const bool want_d = prm.delays != nullptr;
    // values every warp publishes per iteration (constant over the launch: stages that are off in an edge
    // iteration publish identities)
    const unsigned peer_bytes = (unsigned)(CL - 1) * 32u;               // what the peers send for ONE of this CTA's warps

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], kIssuers);
        mbar_init(&xbar[0], NW);
        mbar_init(&xbar[1], NW);
        mbar_fence_init();
    }
    const bool inside = j0 < S;
    if (!inside) {
        // neutral ring tails (the copies only write the live part of the slice)
        const T ninf = from_f32<T>(-INFINITY), zero = from_f32<T>(0.f);
        for (int s = 0; s < NS; ++s) {
            T* sp = reinterpret_cast<T*>(stage0 + (s * PS::kRows) * PS::kRowBytes);
            T* se = reinterpret_cast<T*>(stage0 + (s * PS::kRows + 1) * PS::kRowBytes);
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                sp[l0 + k] = zero;
                if (SOFT) se[l0 + k] = ninf;
            }
        }
    }
    __syncthreads();
    cluster_sync_all();                     // every CTA's exchange barriers exist before anybody stores into them

    const int last = S - 1;
    const bool own_last = mp && (j0 + VPT == S);                        // REPLACE rule: column S-1 is the thread's last element

    const unsigned row_bytes = (unsigned)(slice_n * (int)sizeof(T));
    auto stage_p = [&](int s) { return reinterpret_cast<T*>(stage0 + (s * PS::kRows) * PS::kRowBytes); };
    auto stage_e = [&](int s) { return reinterpret_cast<T*>(stage0 + (s * PS::kRows + 1) * PS::kRowBytes); };
    auto issue = [&](int i, int s) {
        if (!has_data) return;
        if (warp == 0) {
            if (elect_one()) {
                mbar_expect_tx(&bars[s], (SOFT && kIssuers == 1) ? 2u * row_bytes : row_bytes);
                tma_load_1d(stage_p(s), gp + (size_t)i * S + slice_base, row_bytes, &bars[s]);
                if (SOFT && kIssuers == 1) tma_load_1d(stage_e(s), ge + (size_t)i * S + slice_base, row_bytes, &bars[s]);
            }
        } else if (SOFT && kIssuers == 2 && warp == 1) {
            if (elect_one()) {
                mbar_expect_tx(&bars[s], row_bytes);
                tma_load_1d(stage_e(s), ge + (size_t)i * S + slice_base, row_bytes, &bars[s]);
            }
        }
    };
    for (int i = 0; i < NS && i < T_len; ++i) issue(i, i);

    // Lane l < 2 * (CL - 1) of every warp sends one 16-byte half of the warp's eight values to one peer: the
    // whole exchange of a warp is ONE st.async instruction (one lane per message) instead of 2 * (CL - 1)
    // single-lane instructions in a row.
    const bool sender = lane < 2 * (CL - 1);
    const unsigned peer_rank = (rank + 1u + (unsigned)(lane >> 1)) % (unsigned)CL;
    const uint32_t peer_x = mapa_u32(smem_u32(xbuf + gw * kXW + (lane & 1) * 4), peer_rank);
    const uint32_t peer_bar = mapa_u32(smem_u32(&xbar[0]), peer_rank);

    const float one_eps = 1.0f + eps;
    float2 a_prev[H];
    float2 rc[H], P[H], rD[H];
    float2 Rl[H];
#pragma unroll
    for (int q = 0; q < H; ++q) {
        a_prev[q] = make_float2((j0 + 2 * q == 0) ? 1.0f : 0.0f, 0.0f);
        rc[q] = P[q] = rD[q] = Rl[q] = f2(0.f);
    }
    float m_cur = 0.f;
    float rt_prev = 0.f, rsum_prev = 0.f, wsum_prev = 0.f;
    float a_last_raw = 0.f;
    float rdl_pub = 0.f;                    // owner thread: 1/D at the mass-preservation column of the row INV just finished
    float rdl_next = 0.f, rdl_cur = 0.f;    // the same value after the exchange, for the row RECR finishes next / now
    unsigned umax = 0u;
    bool bad = false, nan_out = false;

    int slotI = NS - 1;
    unsigned parI = 1u;
    int sb_w = kExStash - 1;
    int kx = 0;                             // exchange counter: buffer kx & 1, barrier phase (kx >> 1) & 1

    auto body = [&](auto steady_c, const int it) __attribute__((always_inline)) {
        constexpr bool STEADY = decltype(steady_c)::value;
        const bool doM = SOFT && (STEADY || it + 2 < T_len);
        const bool doI = STEADY || (it >= -1 && it + 1 < T_len);
        const bool doU = STEADY || (it >= 0 && it < T_len);
        const bool doR = STEADY || (it >= 1 && it - 1 < T_len);
        const int xb = kx & 1;
        float* xw = xbuf + xb * (kPipeSlots * kXStride);
        const uint32_t xoff = (uint32_t)(xb * (kPipeSlots * kXStride) * 4);
        // ================================================================ PRE
        float em = -INFINITY;
        if (SOFT && doM) {
            int slotM = slotI + 1;
            unsigned parM = parI;
            if (slotM == NS) { slotM = 0; parM ^= 1u; }
            if (has_data) mbar_wait(&bars[slotM], parM);
            float2 Em[H];
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(stage_e(slotM) + l0, Em, dummy);
            em = fmaxf(Em[0].x, Em[0].y);
#pragma unroll
            for (int q = 1; q < H; ++q) em = fmaxf(em, fmaxf(Em[q].x, Em[q].y));
        }
        float2 p_n[H], cpre[H], Dl[H];
        float xinc = 1.f, einc = 0.f;
        if (doI) {
            if (!SOFT && has_data) mbar_wait(&bars[slotI], parI);
            lds_row2<T, VPT, true>(stage_p(slotI) + l0, p_n, umax);
            float2 E_n[H];
            if (SOFT) {
                unsigned dummy = 0u;
                lds_row2<T, VPT, false>(stage_e(slotI) + l0, E_n, dummy);
            }
            xinc = local_cumprod<VPT>(p_n, eps, cpre);
            if (SOFT) {
                float2 unused[H], ex_n[H];
                einc = local_exp_prefix<VPT, false>(E_n, m_cur, eps, unused, ex_n, Dl);
                nan_out = nan_out || (einc != einc);
#pragma unroll
                for (int q = 0; q < VPT / 4; ++q)
                    stash[(sb_w * (VPT / 4) + q) * THREADS + tid] =
                        make_float4(ex_n[2 * q].x, ex_n[2 * q].y, ex_n[2 * q + 1].x, ex_n[2 * q + 1].y);
            }
        }
        float2 sl[H];
        float uinc = 0.f;
        if (doU) {
            uinc = local_u_prefix<VPT>(a_prev, rc, sl);
            nan_out = nan_out || (uinc != uinc);
        }
        float rinc = rt_prev;
        float xexc, eexc, uexc, rexc;
        float wm = -INFINITY;
        if (SOFT) {
            wscan_xeur(xinc, einc, uinc, rinc);
            wneigh_xeur(xinc, einc, uinc, rinc, xexc, eexc, uexc, rexc);
            wm = wmax_redux(em);
        } else {
            wscan_xu(xinc, uinc);
            xexc = wprev(xinc, 1.f);
            uexc = wprev(uinc, 0.f);
            eexc = rexc = 0.f;
        }
        float ws = 0.f, wr = 0.f, wd = 0.f;
        if (mp) {
            ws = warp_sum(rsum_prev);
            if (SOFT) wr = warp_sum(rdl_pub);               // non-zero in one thread of one warp of the row
        }
        if (want_d) wd = warp_sum(wsum_prev);
        // the warp's eight values {row max, x, e, u totals (lane 31), r total (lane 0), row sum, delay sum,
        // 1/D at the mass-preservation column}: into the own exchange buffer and, as two 16-byte stores per
        // peer, into every other CTA's
        {
            const float x31 = __shfl_sync(kFull, xinc, 31), e31 = __shfl_sync(kFull, einc, 31), u31 = __shfl_sync(kFull, uinc, 31);
            const float r0 = __shfl_sync(kFull, rinc, 0);
            const float4 lo = make_float4(wm, x31, e31, u31), hi = make_float4(r0, ws, wd, wr);
            if (lane < 2) reinterpret_cast<float4*>(xw + gw * kXW)[lane] = lane ? hi : lo;
            if (sender) st_async_f32x4(peer_x + xoff, (lane & 1) ? hi : lo, peer_bar + (uint32_t)(xb * 8));
        }
        // ---- the exchange: this warp's values are on their way to every CTA; arrive on the own barrier
        // (release: orders this CTA's shared-memory traffic of the iteration like __syncthreads did) and wait
        // for the own warps and for the peers' bytes
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx_release(&xbar[xb], peer_bytes);
        mbar_wait_cluster(&xbar[xb], (unsigned)((kx >> 1) & 1));
        ++kx;
        if (doI && it + 1 + NS < T_len) issue(it + 1 + NS, slotI);
        if (++slotI == NS) { slotI = 0; parI ^= 1u; }

        // ================================================================ POST
        if (SOFT && doM) m_cur = cw_max<NWC>(xw + 0, lane);
        if (SOFT && mp) {
            rdl_cur = rdl_next;
            rdl_next = cw_sum<NWC>(xw + 7, lane);
        }
        if (doR) {
            const int i = it - 1;
            float resid = 0.f, row_total = 0.f;
            if (mp) {
                row_total = cw_sum<NWC>(xw + 5, lane);
                resid = 1.0f - fminf(fmaxf(row_total, 0.0f), 1.0f);
            }
            if (SOFT) {
                float rbase = cw_suffix_add<NWC>(xw + 4, gw, lane) + rexc;
                if (mp) rbase += resid * rdl_cur;
                const float2 rb = f2(rbase);
                int sb_r = sb_w + 1;
                if (sb_r == kExStash) sb_r = 0;
                float2 b[H];
#pragma unroll
                for (int q = 0; q < VPT / 4; ++q) {
                    const float4 e4 = stash[(sb_r * (VPT / 4) + q) * THREADS + tid];
                    b[2 * q] = min2(mul2(f2(e4.x, e4.y), add2(rb, Rl[2 * q])), 1.0f);
                    b[2 * q + 1] = min2(mul2(f2(e4.z, e4.w), add2(rb, Rl[2 * q + 1])), 1.0f);
                }
                if (inside) st_row2_f32<VPT, true>(g_beta + (size_t)i * S, j0, S, true, b);
            }
            if (want_d) {
                const float wtot = cw_sum<NWC>(xw + 6, lane);
                if (rank == 0 && tid == 0) prm.delays[(size_t)n * T_len + i] = mp ? wtot + (float)(last + 1) * resid : wtot;
            }
            if (own_last) {
                g_alpha[(size_t)i * S + last] = resid;
                if (prm.side != nullptr)
                    *reinterpret_cast<float2*>(prm.side + ((size_t)n * T_len + i) * 2) = make_float2(a_last_raw, row_total);
            }
        }
        if (doU) {
            const float ubase = cw_prefix_add<NWC>(xw + 3, gw, lane) + uexc;
            float2 sfull[H], z[H];
            finish_u_prefix<VPT>(ubase, sl, P, sfull, z);
#pragma unroll
            for (int q = 0; q < H; ++q) a_prev[q] = min2(z[q], 1.0f);
            if (inside) st_row2_f32<VPT, true>(g_alpha + (size_t)it * S, j0, S, true, a_prev);
            if (mp || SOFT || want_d) {
                float2 a_s[H];
#pragma unroll
                for (int q = 0; q < H; ++q) a_s[q] = a_prev[q];
                if (own_last) {
                    a_last_raw = a_s[H - 1].y;
                    a_s[H - 1].y = 0.f;
                }
                if (SOFT) rt_prev = local_r_suffix<VPT>(a_s, rD, Rl);
                if (mp) {
                    float2 acc = a_s[0];
#pragma unroll
                    for (int q = 1; q < H; ++q) acc = add2(acc, a_s[q]);
                    rsum_prev = acc.x + acc.y;
                }
                if (want_d) {
                    const float2 fj = f2((float)j0);
                    float2 acc = mul2(a_s[0], add2(fj, f2(1.0f, 2.0f)));
#pragma unroll
                    for (int q = 1; q < H; ++q) acc = fma2(a_s[q], add2(fj, f2((float)(2 * q + 1), (float)(2 * q + 2))), acc);
                    wsum_prev = acc.x + acc.y;
                }
            }
        }
        if (doI) {
            const float xoffm = cw_prefix_mul<NWC>(xw + 1, gw, lane);
            const float cbase = (one_eps * xoffm) * xexc;
            float2 cp[H];
            finish_cumprod<VPT>(cbase, cpre, p_n, eps, cp, rc, P);
            if (SOFT) {
                const float ebase = cw_prefix_add<NWC>(xw + 2, gw, lane) + eexc;
                finish_exp_prefix<VPT>(ebase, eps, Dl, rD);
                if (own_last) rdl_pub = rD[H - 1].y;
            }
        }
        if (++sb_w == kExStash) sb_w = 0;
    };

    using Steady = std::integral_constant<bool, true>;
    using Edge = std::integral_constant<bool, false>;
    int it = -2;
    for (; it < 1 && it <= T_len; ++it) body(Edge{}, it);
    for (; it <= T_len - 3; ++it) body(Steady{}, it);
    for (; it <= T_len; ++it) body(Edge{}, it);

    if (prm.status != nullptr) {
        if (nan_out) atomicOr(prm.status, SIMULST_ST_NAN);
        bad = umax_trips<T>(umax);
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
    cluster_sync_all();                     // no CTA leaves while a peer may still store into its shared memory
}

// One cluster of CL CTAs per row.  Returns 1 when the configuration does not fit.
template <int THREADS, int VPT, typename T, bool SOFT, int CL>
int launch_mma_fwd_cluster_impl(const MmaParams& prm, cudaStream_t stream) {
    using PS = PipeStatic<THREADS, VPT, T, SOFT>;
    if (!PS::kFits || PS::kNS < 2) return 1;
    auto kern = mma_fwd_cluster_kernel<THREADS, VPT, T, SOFT, CL>;
    static size_t attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (PS::kTotal > attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS::kTotal) != cudaSuccess) {
            cudaGetLastError();
            return SIMULST_E_LAUNCH;
        }
        attr_set[dev & 63] = PS::kTotal;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)prm.N * CL, 1, 1);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = PS::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, prm);
    return check_launch();
}

// Cluster shape for a row of S frames: slices of at most 1024 frames (96 or 128 threads per CTA), i.e. 2 / 4 / 8
// CTAs up to 2048 / 4096 / 8192 frames.  Measured at 64 rows x 128 steps (tests/dev_cluster_probe.py, every shape
// that holds the row): the time depends on the threads per CTA and on the CTAs per row, hardly on S --
// 2 CTAs x 96 / 128 / 192 / 256 threads: 131 / 156 / 195 / 210 us; 4 CTAs: 201 / 219 / 257 / 297 us; 8 CTAs x 96 /
// 128: 244 / 288 us -- and fewer, fatter CTAs never beat the smallest slices that hold the row.
// prm.cluster_cl / cluster_threads (simulst_mma_set_cluster_shape) force a shape.  Returns 1 when the call does
// not qualify (the caller then takes the single-CTA kernels).
template <typename T, bool SOFT>
int launch_mma_fwd_cluster(const MmaParams& prm, cudaStream_t stream) {
    const int S = prm.S;
    if (prm.mask != nullptr || !prm.tma || !prm.vec_out || prm.pitched || S % 8 != 0 || S <= 1024 || S > 8192) return 1;
    int cl, th;
    if (prm.cluster_cl != 0) {
        cl = prm.cluster_cl;
        th = prm.cluster_threads;
    } else {
        cl = S <= 2048 ? 2 : (S <= 4096 ? 4 : 8);
        th = (S + cl - 1) / cl <= 768 ? 96 : 128;           // frames per CTA: 513 .. 1024
    }
    if (cl * th * 8 < S) return 1;
#define SIMULST_CL(CLV, TH) \
    if (cl == CLV && th == TH) return launch_mma_fwd_cluster_impl<TH, 8, T, SOFT, CLV>(prm, stream);
    SIMULST_CL(2, 96) SIMULST_CL(2, 128) SIMULST_CL(4, 96) SIMULST_CL(4, 128) SIMULST_CL(8, 96) SIMULST_CL(8, 128)
#undef SIMULST_CL
    return 1;
}

}  // namespace simulst
