// CIF forward / backward, TMA-staged tile kernels (the fast path of simulst_cif_fwd / _bwd).
//
// Both directions are streaming passes over input[B,S,C] whose only irregularity is the
// data-dependent mapping frame -> output slot.  That mapping is monotone, so a CTA's working set
// is always ONE contiguous run of frames and ONE contiguous run of slot rows:
//   backward: a tile of FR consecutive frames of row b needs grad_out rows l(first)..r(last)
//   forward : a tile of 8 consecutive slots of row b needs frames seg_first[t0]..seg_first[t0+8]
// Each run is fetched with a single 1-D bulk copy (cp.async.bulk -> UBLKCP, completion on an
// mbarrier) issued by one thread, so tens of KB per CTA are in flight without holding registers;
// warps then work out of shared memory (lanes over channels, conflict-free 16-byte accesses) and
// write their result rows straight to HBM with coalesced vector stores.
// Arithmetic (order of every sum included) is identical to the per-warp fallback kernels in
// cif.cu, which remain for rows that are not 16-byte aligned.
#pragma once

#include "common.cuh"

namespace simulst {

constexpr int kTileWarps = 8;
constexpr int kTileThreads = kTileWarps * kWarp;
constexpr int kTileHeader = 128;     // mbarriers + tile metadata

// ---------------------------------------------------------------------------- backward tile
// grid = B * ceil(S / FR); shared memory = header + FR rows of x + GR rows of grad_out.
template <typename TX, typename TA, int NP>
__global__ void __launch_bounds__(kTileThreads)
cif_bwd_tile_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
                    const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                    const TX* __restrict__ g_out, const TX* __restrict__ g_delay,
                    const float* __restrict__ tail_weights, const int64_t* __restrict__ len0,
                    const int64_t* __restrict__ len1,
                    TX* __restrict__ g_x, float* __restrict__ ws_gl, float* __restrict__ ws_gd,
                    int B, int S, int C, int T, int T_out, float beta, float tail_thres, int training,
                    int FR, int GR) {
    constexpr int V = 4;
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);          // [0] x tile, [1] grad_out tile
    int* meta = reinterpret_cast<int*>(smem + 32);               // [0] first staged slot, [1] rows staged
    TX* xt = reinterpret_cast<TX*>(smem + kTileHeader);
    TX* gt = xt + (size_t)FR * C;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = (S + FR - 1) / FR;
    const int b = blockIdx.x / tiles;
    const int s0 = (blockIdx.x % tiles) * FR;
    const int nf = min(FR, S - s0);
    const float* cs = csum + (size_t)b * S;
    const size_t row_bytes = (size_t)C * sizeof(TX);

    long long l0 = 0, l1 = 0;
    float tail_mul = 1.0f;
    if (!training) {
        l0 = len0[b]; l1 = len1[b];
        if (l1 > l0) tail_mul = __fdiv_rn(beta, tail_weights[b]);
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        mbar_expect_tx(&bars[0], (unsigned)(nf * row_bytes));
        tma_load_1d(xt, x + ((size_t)b * S + s0) * C, (unsigned)(nf * row_bytes), &bars[0]);
        // slots fed by this tile: l(first frame) .. r(last frame); rows that were sliced off or
        // zeroed never contribute
        const int t_first = (s0 == 0) ? 0 : fire_index(cs[s0 - 1], beta, T);
        const int t_last = fire_index(cs[s0 + nf - 1], beta, T);
        long long lim = T_out;
        if (!training && l1 < lim) lim = l1;
        int n = (int)min((long long)t_last, lim - 1) - t_first + 1;
        n = max(0, min(n, GR));
        meta[0] = t_first;
        meta[1] = n;
        if (n > 0) {
            mbar_expect_tx(&bars[1], (unsigned)(n * row_bytes));
            tma_load_1d(gt, g_out + ((size_t)b * T_out + t_first) * C, (unsigned)(n * row_bytes), &bars[1]);
        }
    }
    // ---- per-frame scalars of this warp's frames (frame f = warp + 8 i is held by lane i)
    float c_s = 0.f, c_p = 0.f, a_w = 0.f;
    {
        const int f = warp + kTileWarps * lane;
        if (f < nf) {
            const int s = s0 + f;
            c_s = cs[s];
            c_p = (s == 0) ? 0.f : cs[s - 1];
            const bool pad = mask && mask[(size_t)b * S + s];
            a_w = pad ? 0.f : to_f32<TA>(alpha[(size_t)b * S + s]) * scale[b];
        }
    }
    __syncthreads();                                   // barriers initialised, meta visible
    const int t_base = meta[0], n_stage = meta[1];
    mbar_wait(&bars[0], 0u);
    if (n_stage > 0) mbar_wait(&bars[1], 0u);

    auto live = [&](int t) -> bool { return t < T_out && (training || t < l1); };
    auto factor = [&](int t) -> float { return (!training && t == l0 && l1 > l0) ? tail_mul : 1.0f; };
    // grad_out row of slot t: staged copy if it is in the tile, else global memory
    auto g_row = [&](int t) -> const TX* {
        const int k = t - t_base;
        return (k >= 0 && k < n_stage) ? gt + (size_t)k * C : g_out + ((size_t)b * T_out + t) * C;
    };

    for (int i = 0; warp + kTileWarps * i < nf; ++i) {
        const int f = warp + kTileWarps * i;
        const int s = s0 + f;
        const float cs_s = __shfl_sync(kFull, c_s, i);
        const float cs_p = __shfl_sync(kFull, c_p, i);
        const float a = __shfl_sync(kFull, a_w, i);
        const int r = fire_index(cs_s, beta, T);
        const int l = (s == 0) ? 0 : fire_index(cs_p, beta, T);
        const float pos = (float)(s + 1);
        const TX* xs = xt + (size_t)f * C;
        TX* gx = g_x + ((size_t)b * S + s) * C;
        const bool use_l = live(l), use_r = r > l && live(r);
        float dot_l = 0.f, dot_r = 0.f;
        for (int c0 = 0; c0 < C; c0 += kWarp * V * NP) {
            float xv[NP][V], acc[NP][V];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c = c0 + (q * kWarp + lane) * V;
#pragma unroll
                for (int k = 0; k < V; ++k) { xv[q][k] = 0.f; acc[q][k] = 0.f; }
                if (c < C) load_vec<TX, V>(xs + c, C - c, true, 0.f, xv[q]);
            }
            auto apply = [&](int t) {
                const float fct = factor(t);
                const float w = slot_weight(t, l, r, a, cs_s, beta);
                const TX* gr = g_row(t);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int c = c0 + (q * kWarp + lane) * V;
                    if (c < C) {
                        float gv[V];
                        load_vec<TX, V>(gr + c, C - c, true, 0.f, gv);
                        float d = 0.f;
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            acc[q][k] += (w * fct) * gv[k];
                            d += gv[k] * xv[q][k];
                        }
                        if (t == l) dot_l += d * fct;
                        if (t == r) dot_r += d * fct;
                    }
                }
            };
            if (use_l) apply(l);
            for (int t = l + 1; t < r; ++t) {               // frames that fire more than once
                if (!live(t)) { if (t >= T_out) break; continue; }
                apply(t);
            }
            if (use_r) apply(r);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c = c0 + (q * kWarp + lane) * V;
                if (c < C) store_vec<TX, V>(gx + c, C - c, true, acc[q]);
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
}

// ---------------------------------------------------------------------------- forward tile
// grid = B * ceil(T_alloc / 8): one warp per output slot, 8 consecutive slots per CTA.  The
// frames the 8 slots draw from are staged FC at a time.
template <typename TX, typename TA, int NP>
__global__ void __launch_bounds__(kTileThreads)
cif_fwd_tile_kernel(const TX* __restrict__ x, const float* __restrict__ csum, const float* __restrict__ scale,
                    const TA* __restrict__ alpha, const uint8_t* __restrict__ mask,
                    const int* __restrict__ seg_first, int seg_stride,
                    TX* __restrict__ out, TX* __restrict__ delays, float* __restrict__ tail_weights,
                    const int64_t* __restrict__ lengths, int64_t* __restrict__ lengths_out,
                    int* __restrict__ t_max2,
                    int B, int S, int C, int T, int T_alloc, float beta, float tail_thres, int training,
                    int FC) {
    constexpr int V = 4;
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    TX* xt = reinterpret_cast<TX*>(smem + kTileHeader);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = (T_alloc + kTileWarps - 1) / kTileWarps;
    const int b = blockIdx.x / tiles;
    const int t0 = (blockIdx.x % tiles) * kTileWarps;
    const int t = t0 + warp;
    const bool have = t < T_alloc;
    const float* cs = csum + (size_t)b * S;
    const TA* a_row = alpha + (size_t)b * S;
    const uint8_t* m_row = mask ? mask + (size_t)b * S : nullptr;
    const int* f_row = seg_first + (size_t)b * seg_stride;
    const float sc = scale[b];
    const size_t row_bytes = (size_t)C * sizeof(TX);

    auto range_hi = [&](int tt) -> int {
        return (tt >= T || tt + 1 >= seg_stride) ? S - 1 : min(f_row[tt + 1], S - 1);
    };
    // this warp's frames, and the CTA's (slot ranges are monotone, so first slot .. last slot)
    int s_lo = 1, s_hi = 0;
    if (have) {
        s_lo = (t < seg_stride) ? f_row[t] : S;
        s_hi = range_hi(t);
    }
    const int t_last = min(t0 + kTileWarps, T_alloc) - 1;
    const int cta_lo = (t0 < seg_stride) ? f_row[t0] : S;
    const int cta_hi = range_hi(t_last);

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    float acc[NP][V];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[q][k] = 0.f;
    float ws = 0.f, ds = 0.f;
    unsigned parity = 0u;
    for (int cb = cta_lo; cb <= cta_hi; cb += FC) {
        const int nfc = min(FC, cta_hi - cb + 1);
        if (tid == 0) {
            mbar_expect_tx(bar, (unsigned)(nfc * row_bytes));
            tma_load_1d(xt, x + ((size_t)b * S + cb) * C, (unsigned)(nfc * row_bytes), bar);
        }
        const int a_lo = max(s_lo, cb), a_hi = min(s_hi, cb + nfc - 1);
        // per-frame weights of the first 32 frames of this warp's share (overlaps the copy)
        float w = 0.f, d = 0.f;
        auto weights = [&](int base) {
            const int s = base + lane;
            w = 0.f; d = 0.f;
            if (s <= a_hi) {
                const float c_s = cs[s];
                const int r = fire_index(c_s, beta, T);
                const int l = (s == 0) ? 0 : fire_index(cs[s - 1], beta, T);
                const float a = (m_row && m_row[s]) ? 0.f : to_f32<TA>(a_row[s]) * sc;
                if (t >= l && t <= r) {
                    w = slot_weight(t, l, r, a, c_s, beta);
                    d = (l != r && t != l && t != r) ? (float)(s + 1) : __fdiv_rn(w * (float)(s + 1), beta);
                }
            }
        };
        if (a_lo <= a_hi) weights(a_lo);
        mbar_wait(bar, parity);
        parity ^= 1u;
        for (int base = a_lo; base <= a_hi; base += kWarp) {
            if (base != a_lo) weights(base);
            const int n = min(kWarp, a_hi - base + 1);
            for (int f = 0; f < n; ++f) {
                const float wu = __shfl_sync(kFull, w, f);
                const float du = __shfl_sync(kFull, d, f);
                ws += wu;
                ds += du;
                const TX* xs = xt + (size_t)(base + f - cb) * C;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int c = (q * kWarp + lane) * V;
                    if (c < C) {
                        float xv[V];
                        load_vec<TX, V>(xs + c, C - c, true, 0.f, xv);
#pragma unroll
                        for (int k = 0; k < V; ++k) acc[q][k] += wu * xv[k];
                    }
                }
            }
        }
        if (cb + FC <= cta_hi) __syncthreads();          // the buffer is refilled by the next chunk
    }
    if (!have) return;
    // tail handling (inference): the slot at the row's own length holds the partial segment
    float mul = 1.0f;
    bool zero = false;
    if (!training) {
        const long long len0 = lengths[b];
        if (t == len0) {
            if (ws >= tail_thres) mul = __fdiv_rn(beta, ws); else zero = true;
        } else if (t > len0) {
            zero = true;
        }
    }
    TX* o_row = out + ((size_t)b * T_alloc + t) * C;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int c = (q * kWarp + lane) * V;
        if (c < C) {
            float ov[V];
#pragma unroll
            for (int k = 0; k < V; ++k) ov[k] = zero ? 0.f : acc[q][k] * mul;
            store_vec<TX, V>(o_row + c, C - c, true, ov);
        }
    }
    if (lane == 0) {
        delays[(size_t)b * T_alloc + t] = from_f32<TX>(ds);
        if (!training && t == lengths[b]) {
            tail_weights[b] = ws;                        // single writer per row
            const long long len1 = lengths[b] + (ws >= tail_thres ? 1 : 0);
            lengths_out[b] = len1;
            atomicMax(t_max2, (int)len1);
        }
    }
}

}  // namespace simulst
