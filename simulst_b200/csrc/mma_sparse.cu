// Fixed pre-decision ("pooled p_choose") training path on the POOLED GRID.
//
// The reference's *_fixed_pre_decision classes (codebase/modules/fixed_pre_decision.py:85-95,
// :139-159) zero-upsample p_choose_pooled [N,T,Sp] to [N,T,S]: dense column j carries
// pooled[(j+1)/r - 1] when (j+1) % r == 0, column S-1 carries pooled[Sp-1], every other column
// is zero.  Everything the expected-alignment recurrence produces is then zero off that grid:
//     alpha_ij = clamp(p_ij * cp_ij * s_ij, 0, 1) = 0          wherever p_ij = 0,
// and its backward needs nothing from off-grid columns (their gradients only reach p = 0 inputs
// that have no producer).  Only two things stay dense: the expected soft attention
// (codebase/utils/monotonic_attention.py:79-152), whose rows do not depend on each other at all,
// and the dense [N,T,S] alpha / beta tensors the module hands to its callers.  So instead of one
// CTA per (batch, head) row walking T steps over S columns (the dense kernels), this path runs
//
//   K1 sparse_alpha_fwd   one CTA per (batch, head) row, T sequential steps over Sp = S/r
//                         columns: cumprod as a product scan with the (1+eps) factors of the
//                         zero columns folded into per-column constants, prefix sum, clamp, mass
//                         preservation, expected delays.  Writes alpha on the grid [N,T,Sp]
//                         (+ the residual of a right-padded row when it lands off the grid).
//   K2 sparse_row_fwd     one CTA per block of (n,t) rows, no sequential dependency: energy
//                         rows by TMA ring, exp / prefix scan / suffix scan, beta out, dense alpha
//                         (and optionally the dense p_choose) written from the grid values.
//   K3 sparse_row_bwd     the same rows backwards: grad_beta, energy (and grad_alpha) rows by
//                         TMA, grad_energy out, d/d alpha collected ON THE GRID [N,T,Sp].
//   K4 sparse_alpha_bwd   K1 backwards on the grid; grad of the pooled p_choose out.
//
// Bytes per dense element (bf16 in, ratio 8): forward 2 (energy) + 8 (alpha, beta) + ~1 (grid
// traffic) instead of 12; backward 2 + 4 + [4] + 2 + ~1.5 instead of 20 -- and the two streaming
// kernels are plain bandwidth-bound row kernels instead of T-deep latency chains.
//
// Formulas: SURVEY Appendix A.1-A.4 (verified against the reference's autograd).  Masks: none, or a
// right-padding mask promised by SIMULST_MMA_RIGHT_PADDING (verified per row, like the dense
// MASKED kernels: a violation sets SIMULST_ST_NOT_RIGHT_PADDED and poisons the row with NaN).
#include <algorithm>
#include <cstdlib>

#include "mma_scan.cuh"
#include "mma_sparse.h"

namespace simulst {

namespace {

// development knob (SIMULST_SPARSE_VARIANT): 0 warp-specialised K1/K4 (default), 2 generic block-scan kernels
static const int g_sparse_variant = [] { const char* e = getenv("SIMULST_SPARSE_VARIANT"); return e ? atoi(e) : 0; }();

constexpr float kLog2eS = 1.4426950408889634f;
// K2 / K3: rows staged ahead by TMA (ring depth; long rows get a shallower ring to fit 227 KB)
__host__ __device__ constexpr int ring_fwd(int cap) { return cap <= 4096 ? 4 : 2; }
__host__ __device__ constexpr int ring_bwd(int cap) { return cap <= 4096 ? 3 : 2; }

__device__ __forceinline__ float ex2a(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ bool elect1() { return elect_one(); }

// VPT consecutive elements of type T moved in 16-byte chunks (rows are 16-byte aligned, not more)
template <typename T, int VPT>
__device__ __forceinline__ void load_typed(const T* __restrict__ src, float (&v)[VPT]) {
    constexpr int PK = 16 / (int)sizeof(T);
#pragma unroll
    for (int c = 0; c < VPT / PK; ++c) {
        const Pack<T, PK> pk = *reinterpret_cast<const Pack<T, PK>*>(src + c * PK);
#pragma unroll
        for (int k = 0; k < PK; ++k) v[c * PK + k] = to_f32<T>(pk.v[k]);
    }
}
template <typename T, int VPT>
__device__ __forceinline__ void store_typed(T* __restrict__ dst, int n_valid, const T (&v)[VPT]) {
    constexpr int PK = 16 / (int)sizeof(T);
    if (n_valid >= VPT) {
#pragma unroll
        for (int c = 0; c < VPT / PK; ++c) {
            Pack<T, PK> pk;
#pragma unroll
            for (int k = 0; k < PK; ++k) pk.v[k] = v[c * PK + k];
            *reinterpret_cast<Pack<T, PK>*>(dst + c * PK) = pk;
        }
    } else {
#pragma unroll
        for (int k = 0; k < VPT; ++k)
            if (k < n_valid) dst[k] = v[k];
    }
}

// dense column of pooled element m / pooled element of dense column j (-1: off the grid)
__device__ __forceinline__ int grid_col(int m, int Sp, int S, int r) { return m == Sp - 1 ? S - 1 : (m + 1) * r - 1; }
__device__ __forceinline__ int grid_idx(int j, int Sp, int S, int r) {
    if (j == S - 1) return Sp - 1;
    return ((j + 1) % r == 0) ? (j + 1) / r - 1 : -1;
}

// ---- block-wide scans of one value per thread.  NW == 1: shuffles only.  Otherwise one value
// per warp through `slot` (NW floats, owned by this exchange until the block's next use of it two
// exchanges later) and ONE __syncthreads.
template <int NW>
__device__ __forceinline__ float2 block_prefix(float v, float* slot, int warp, int lane) {   // {exclusive, -}
    const float inc = wscan_prefix_add(v);
    const float exc = wprev(inc, 0.f);
    if constexpr (NW == 1) {
        return make_float2(exc, 0.f);
    } else {
        if (lane == 31) slot[warp] = inc;
        __syncthreads();
        return make_float2(xw_prefix_add<NW>(slot, warp, lane) + exc, 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float2 block_suffix(float v, float* slot, int warp, int lane) {   // {exclusive, -}
    const float inc = wscan_suffix_add(v);
    const float exc = wnext(inc, 0.f);
    if constexpr (NW == 1) {
        return make_float2(exc, 0.f);
    } else {
        if (lane == 0) slot[warp] = inc;
        __syncthreads();
        return make_float2(xw_suffix_add<NW>(slot, warp, lane) + exc, 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float block_prefix_mul(float v, float* slot, int warp, int lane) {   // exclusive product
    const float inc = wscan_prefix_mul(v);
    const float exc = wprev(inc, 1.0f);
    if constexpr (NW == 1) {
        return exc;
    } else {
        if (lane == 31) slot[warp] = inc;
        __syncthreads();
        return xw_prefix_mul<NW>(slot, warp, lane) * exc;
    }
}
template <int NW>
__device__ __forceinline__ float block_max(float v, float* slot, int warp, int lane) {
    const float m = wmax_redux(v);
    if constexpr (NW == 1) {
        return m;
    } else {
        if (lane == 0) slot[warp] = m;
        __syncthreads();
        return xw_max<NW>(slot, lane);
    }
}
template <int NW>
__device__ __forceinline__ float block_sum1(float a, float* slot, int warp, int lane) {
    a = warp_sum(a);
    if constexpr (NW > 1) {
        if (lane == 0) slot[warp] = a;
        __syncthreads();
        a = combine_sum<NW>(slot, lane);
    }
    return a;
}
// three sums at once (one barrier)
template <int NW>
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float* slot, int warp, int lane) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if constexpr (NW > 1) {
        if (lane == 0) { slot[warp] = a; slot[32 + warp] = b; slot[64 + warp] = c; }
        __syncthreads();
        a = combine_sum<NW>(slot, lane); b = combine_sum<NW>(slot + 32, lane); c = combine_sum<NW>(slot + 64, lane);
    }
}

// Exchange slots: kXSlots exchanges per iteration, 96 floats each, double-buffered by iteration.
constexpr int kXFloats = 96;
template <int SLOTS>
struct Xs {
    float* base;
    int par;
    __device__ __forceinline__ float* operator()(int s) const { return base + (par * SLOTS + s) * kXFloats; }
    __device__ __forceinline__ void flip() { par ^= 1; }
};

// ---------------------------------------------------------------------------------------------
// Row geometry shared by K1 / K4: live length L of row n (S without a mask), verification of the
// right-padding promise, and the mass-preservation column.
struct RowGeom {
    int L;          // live length; -1: the mask is not a right-padding mask (poisoned row)
    int mp_m;       // pooled index the residual lands on, -1: none (off the grid or no live column)
    int xcol;       // off-grid dense column the residual lands on, -1: none
};
template <int THREADS>
__device__ __forceinline__ RowGeom row_geometry(const SparseParams& prm, int n, int* sh_int) {
    RowGeom g;
    const int S = prm.S, Sp = prm.Sp, r = prm.r;
    if (prm.mask == nullptr) {
        g.L = S; g.mp_m = Sp - 1; g.xcol = -1;
        return g;
    }
    const uint8_t* mrow = prm.mask + (size_t)n * S;
    int live = 0, bad = 0;
    for (int j = threadIdx.x; j < S; j += THREADS) {
        const int dead = mrow[j] != 0;
        live += dead ? 0 : 1;
        if (j + 1 < S && dead && mrow[j + 1] == 0) bad = 1;
    }
    if (threadIdx.x == 0) { sh_int[0] = 0; sh_int[1] = 0; }
    __syncthreads();
    live = __reduce_add_sync(kFull, live);
    bad = __reduce_or_sync(kFull, bad);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh_int[0], live); atomicOr(&sh_int[1], bad); }
    __syncthreads();
    const int L = sh_int[0];
    g.L = sh_int[1] ? -1 : L;
    g.mp_m = -1; g.xcol = -1;
    if (g.L > 0) {
        const int m = grid_idx(L - 1, Sp, S, r);
        if (m >= 0) g.mp_m = m; else g.xcol = L - 1;
    }
    __syncthreads();
    return g;
}

// =============================================================================================
// K1: expected alignment on the grid, forward.  One CTA per (batch, head) row; thread t owns the
// pooled elements [t*EPT, t*EPT + EPT).
template <int NW, int EPT, typename T>
__global__ void __launch_bounds__(NW * 32) sparse_alpha_fwd_kernel(const SparseParams prm) {
    constexpr int THREADS = NW * 32;
    __shared__ float xraw[2 * 3 * kXFloats];
    __shared__ int sh_int[2];
    Xs<3> xs{xraw, 0};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool replace = prm.mask == nullptr;      // mass preservation REPLACES column S-1 (no mask) or ADDS at L-1

    const RowGeom geo = row_geometry<THREADS>(prm, n, sh_int);
    if (tid == 0) { prm.lens[n] = geo.L; prm.xcol[n] = geo.xcol; }
    float* asp = prm.a_sp + (size_t)n * T_len * Sp;
    float* ax = prm.a_x + (size_t)n * T_len;
    if (geo.L < 0) {
        // broken promise: flag it and poison the grid values (K2 poisons the dense outputs)
        if (tid == 0 && prm.status != nullptr) atomicOr(prm.status, SIMULST_ST_NOT_RIGHT_PADDED);
        const float qnan = __int_as_float(0x7fc00000);
        for (size_t q = tid; q < (size_t)T_len * Sp; q += THREADS) asp[q] = qnan;
        for (int q = tid; q < T_len; q += THREADS) ax[q] = qnan;
        return;
    }
    const int L = geo.L;

    // per-element constants
    const int m0 = tid * EPT;
    float W[EPT], wcol[EPT];
    bool valid[EPT], live[EPT];
    const double log1e = log((double)(1.0f + eps));
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int m = m0 + k;
        valid[k] = m < Sp;
        const int col = valid[k] ? grid_col(m, Sp, S, r) : 0;
        live[k] = valid[k] && col < L;
        // cumprod factor of the zero columns in front of this grid column and of the leading
        // ones column (functions.py:28-33): (1+eps)^(1 + col - m)
        W[k] = valid[k] ? (float)exp((double)(1 + col - m) * log1e) : 1.0f;
        wcol[k] = (float)(col + 1);
    }
    const int k_mp = (mp && geo.mp_m >= m0 && geo.mp_m < m0 + EPT) ? geo.mp_m - m0 : -1;
    const float w_x = (float)(geo.xcol + 1);

    const T* gpp = reinterpret_cast<const T*>(prm.pp) + (size_t)n * T_len * Sp;
    float a_prev[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) a_prev[k] = (S == 1 && m0 + k == 0) ? 1.0f : 0.0f;      // alpha_0 = one-hot(0)
    unsigned bits = 0u;
    bool nan_out = false;

    unsigned raw[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) raw[k] = (valid[k] && T_len > 0) ? ldg_raw<T>(gpp + m0 + k) : 0u;

    for (int i = 0; i < T_len; ++i) {
        float p[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const float v = raw_to_f32<T>(raw[k]);
            if (valid[k]) bits |= prob_bits(v) | ((((1.0f - v) + eps) < 0.f) ? SIMULST_ST_NEGPROD : 0u);
            p[k] = live[k] ? v : 0.f;
        }
        if (i + 1 < T_len) {
            const T* nxt = gpp + (size_t)(i + 1) * Sp + m0;
#pragma unroll
            for (int k = 0; k < EPT; ++k) raw[k] = valid[k] ? ldg_raw<T>(nxt + k) : 0u;
        }
        // exclusive cumprod of (1-p)+eps over the dense row, evaluated on the grid
        float xe[EPT], xt = 1.0f;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            xe[k] = xt;
            xt *= valid[k] ? ((1.0f - p[k]) + eps) : 1.0f;
        }
        const float xoff = block_prefix_mul<NW>(xt, xs(0), warp, lane);
        float P[EPT], u[EPT], ut = 0.f;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const float cp = W[k] * (xoff * xe[k]);
            const float c = fminf(fmaxf(cp, eps), 1.0f);
            P[k] = p[k] * cp;
            u[k] = a_prev[k] * fast_rcp(c);
            ut += u[k];
            u[k] = ut;                              // local inclusive prefix
        }
        const float2 up = block_prefix<NW>(ut, xs(1), warp, lane);
        // the one-hot alpha_0 sits on column 0, off the grid unless S == 1: u there is 1/clamp(1+eps) = 1
        const float s_off = up.x + ((i == 0 && S != 1) ? 1.0f : 0.0f);
        float a[EPT], sum_a = 0.f, sum_w = 0.f, sum_x = 0.f;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const float z = P[k] * (s_off + u[k]);
            nan_out = nan_out || (z != z);
            a[k] = fminf(fmaxf(z, 0.0f), 1.0f);
            a_prev[k] = a[k];
            const bool excl = replace && mp && k == k_mp;          // REPLACE: the residual excludes the column itself
            sum_a += excl ? 0.f : a[k];
            sum_w += excl ? 0.f : wcol[k] * a[k];
        }
        block_sum3<NW>(sum_a, sum_w, sum_x, xs(2), warp, lane);
        float res = 0.f;
        if (mp && (geo.mp_m >= 0 || geo.xcol >= 0)) res = 1.0f - fminf(fmaxf(sum_a, 0.0f), 1.0f);
        if (k_mp >= 0) {
#pragma unroll
            for (int k = 0; k < EPT; ++k)
                if (k == k_mp) {
                    if (prm.side != nullptr) prm.side[((size_t)n * T_len + i) * 2] = a[k];
                    a[k] = replace ? res : a[k] + res;
                }
        }
        if (tid == 0) {
            if (mp && prm.side != nullptr) {
                prm.side[((size_t)n * T_len + i) * 2 + 1] = sum_a;
                if (geo.mp_m < 0) prm.side[((size_t)n * T_len + i) * 2] = 0.f;
            }
            ax[i] = (mp && geo.xcol >= 0) ? res : 0.f;
            if (prm.delays != nullptr) {
                float d = sum_w;
                if (mp && geo.mp_m >= 0) d += (float)(grid_col(geo.mp_m, Sp, S, r) + 1) * res;
                if (mp && geo.xcol >= 0) d += w_x * res;
                prm.delays[(size_t)n * T_len + i] = d;
            }
        }
        float* arow = asp + (size_t)i * Sp + m0;
#pragma unroll
        for (int k = 0; k < EPT; ++k)
            if (valid[k]) arow[k] = a[k];
        xs.flip();
    }
    if (prm.status != nullptr) {
        if (nan_out) bits |= SIMULST_ST_NAN;
        bits = __reduce_or_sync(kFull, bits);
        if (lane == 0 && bits) atomicOr(prm.status, bits);
    }
}

// =============================================================================================
// Column pattern of a K2 / K3 thread (row independent): which of its VPT columns lie on the grid,
// and the pooled index of the first of them.
template <int VPT>
struct ColPattern {
    unsigned gbits;     // bit k: column j0+k is on the grid (and < S)
    int m_first;        // pooled index of the lowest on-grid column of this thread
    __device__ __forceinline__ ColPattern(int j0, int S, int Sp, int r) {
        gbits = 0u; m_first = 0;
        bool first = true;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = j0 + k;
            const int m = j < S ? grid_idx(j, Sp, S, r) : -1;
            if (m >= 0) {
                gbits |= 1u << k;
                if (first) { m_first = m; first = false; }
            }
        }
    }
    __device__ __forceinline__ bool on(int k) const { return (gbits >> k) & 1u; }
    __device__ __forceinline__ int idx(int k) const { return m_first + __popc(gbits & ((1u << k) - 1u)); }
};

// =============================================================================================
// K2: rows (n,t) forward.  SOFT: beta = expected soft attention (infinite lookback) from the
// grid alpha and the energy row; always: dense alpha row (and optionally dense p_choose) out.
template <int THREADS, int VPT, typename T, bool SOFT>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 1024 / THREADS : 1)) sparse_row_fwd_kernel(const SparseParams prm, int rows_per_cta) {
    constexpr int NW = THREADS / 32, CAP = THREADS * VPT;
    constexpr int kRing = ring_fwd(CAP);
    constexpr int kRowBytes = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    unsigned char* ring = smem + 128 + 2 * 3 * kXFloats * 4;
    Xs<3> xs{xraw, 0};

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const long long rows_total = (long long)prm.N * T_len;
    const long long row0 = (long long)blockIdx.x * rows_per_cta;
    const int nrows = (int)min((long long)rows_per_cta, rows_total - row0);
    if (nrows <= 0) return;
    const int j0 = tid * VPT;
    const ColPattern<VPT> pat(j0, S, Sp, r);
    const unsigned row_bytes = (unsigned)(S * sizeof(T));

    if (SOFT && tid == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (SOFT) __syncthreads();
    auto issue = [&](int q) {
        if (SOFT && warp == 0 && q < nrows) {
            if (elect1()) {
                uint64_t* bar = &bars[q % kRing];
                mbar_expect_tx(bar, row_bytes);
                tma_load_1d(ring + (q % kRing) * kRowBytes, reinterpret_cast<const T*>(prm.e) + (size_t)(row0 + q) * S,
                            row_bytes, bar);
            }
        }
    };
    for (int q = 0; q < kRing - 1; ++q) issue(q);

    const float qnan = __int_as_float(0x7fc00000);
    // The grid values of a row (alpha on the grid, residual, row geometry) are fetched one row
    // ahead: their global-load latency would otherwise open every row.
    float a_n[VPT], ax_n = 0.f;
    int L_n = 0, xcol_n = -1;
    auto fetch = [&](int q) {
        const long long row = row0 + q;
        const int n = (int)((unsigned)row / (unsigned)T_len);          // N*T < 2^31 (checked by the launcher)
        L_n = __ldg(prm.lens + n);
        xcol_n = __ldg(prm.xcol + n);
        ax_n = __ldg(prm.a_x + row);
        const float* asp = prm.a_sp + (size_t)row * Sp;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            a_n[k] = 0.f;
            if (pat.on(k)) a_n[k] = __ldg(asp + pat.idx(k));
        }
    };
    fetch(0);
    for (int q = 0; q < nrows; ++q) {
        const long long row = row0 + q;
        issue(q + kRing - 1);           // its slot was read in iteration q-1, three barriers ago
        const int L = L_n;
        const int xcol = xcol_n;
        float* arow_out = prm.alpha + (size_t)row * S;
        float* brow_out = SOFT ? prm.beta + (size_t)row * S : nullptr;
        // ---- grid values of this thread's columns
        float a[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) a[k] = (j0 + k == xcol) ? ax_n : a_n[k];
        if (q + 1 < nrows) fetch(q + 1);
        if (prm.p_dense != nullptr && j0 < S) {
            const T* pp = reinterpret_cast<const T*>(prm.pp) + (size_t)row * Sp;
            T* pd = reinterpret_cast<T*>(prm.p_dense) + (size_t)row * S + j0;
            T pv[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) pv[k] = pat.on(k) ? __ldg(pp + pat.idx(k)) : from_f32<T>(0.f);
            store_typed<T, VPT>(pd, S - j0, pv);
        }
        if (L < 0) {            // poisoned row (broken right-padding promise)
#pragma unroll
            for (int k = 0; k < VPT; ++k) a[k] = qnan;
        }
        auto store_row = [&](float* out, const float (&v)[VPT]) {
            if (j0 + VPT <= S) {
#pragma unroll
                for (int c = 0; c < VPT / 4; ++c)
                    *reinterpret_cast<float4*>(out + j0 + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (j0 + k < S) out[j0 + k] = v[k];
            }
        };
        if (prm.alpha != nullptr) store_row(arow_out, a);
        if constexpr (SOFT) {
            mbar_wait(&bars[q % kRing], (unsigned)((q / kRing) & 1));
            const T* erow = reinterpret_cast<const T*>(ring + (q % kRing) * kRowBytes) + j0;
            float E[VPT];
            load_typed<T, VPT>(erow, E);
#pragma unroll
            for (int k = 0; k < VPT; ++k) E[k] = (j0 + k < L) ? E[k] : -INFINITY;
            float mx = E[0];
#pragma unroll
            for (int k = 1; k < VPT; ++k) mx = fmaxf(mx, E[k]);
            const float m = block_max<NW>(mx, xs(0), warp, lane);
            float e[VPT], D[VPT], et = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                e[k] = (j0 + k < L) ? ex2a((E[k] - m) * kLog2eS) + eps : 0.f;
                et += e[k];
                D[k] = et;
            }
            const float2 ep = block_prefix<NW>(et, xs(1), warp, lane);
            float R[VPT], rt = 0.f;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k) {
                // r = alpha / (eps + cumsum(e)); alpha is zero off the grid (and off the residual column)
                const bool sparse_col = pat.on(k) || (j0 + k == xcol);
                if (sparse_col && j0 + k < L) rt += a[k] * fast_rcp(eps + (ep.x + D[k]));
                R[k] = rt;
            }
            const float2 rp = block_suffix<NW>(rt, xs(2), warp, lane);
            float b[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const float v = e[k] * (rp.x + R[k]);
                b[k] = (j0 + k < L) ? fminf(fmaxf(v, 0.0f), 1.0f) : 0.f;
                if (L < 0) b[k] = qnan;
            }
            store_row(brow_out, b);
            xs.flip();
        }
    }
}

// =============================================================================================
// K3: rows (n,t) backward.  SOFT: expected soft attention backward (A.3): grad_energy out,
// d/d alpha' at the grid columns (and at the off-grid residual column) added to the external
// grad_alpha / grad_expected_delays there.  !SOFT: only that gather.
template <int THREADS, int VPT, typename T, bool SOFT>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 768 / THREADS : 1)) sparse_row_bwd_kernel(const SparseParams prm, int rows_per_cta) {
    constexpr int NW = THREADS / 32, CAP = THREADS * VPT;
    constexpr int kRing = ring_bwd(CAP);
    constexpr int kTRow = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    constexpr int kFRow = CAP * 4;
    constexpr int kStage = (SOFT ? kTRow + kFRow : 0) + kFRow;          // energy, grad_beta, grad_alpha
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    int* sh_int = reinterpret_cast<int*>(smem + 128 + 2 * 6 * kXFloats * 4);
    unsigned char* ring = smem + 128 + 2 * 6 * kXFloats * 4 + 256;
    Xs<6> xs{xraw, 0};

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const long long rows_total = (long long)prm.N * T_len;
    const long long row0 = (long long)blockIdx.x * rows_per_cta;
    const int nrows = (int)min((long long)rows_per_cta, rows_total - row0);
    if (nrows <= 0) return;
    const int j0 = tid * VPT;
    const ColPattern<VPT> pat(j0, S, Sp, r);
    const unsigned t_bytes = (unsigned)(S * sizeof(T)), f_bytes = (unsigned)(S * 4);
    const bool has_ga = prm.g_alpha != nullptr;
    const bool has_gb = SOFT && prm.g_beta != nullptr;
    const bool has_gd = prm.g_delays != nullptr;
    const bool any_tma = SOFT || has_ga;

    if (any_tma && tid == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int q) {
        if (any_tma && warp == 0 && q < nrows) {
            if (elect1()) {
                uint64_t* bar = &bars[q % kRing];
                unsigned char* st = ring + (q % kRing) * kStage;
                const size_t ro = (size_t)(row0 + q) * S;
                mbar_expect_tx(bar, (SOFT ? t_bytes : 0u) + (has_gb ? f_bytes : 0u) + (has_ga ? f_bytes : 0u));
                if (SOFT) tma_load_1d(st, reinterpret_cast<const T*>(prm.e) + ro, t_bytes, bar);
                if (has_gb) tma_load_1d(st + kTRow, prm.g_beta + ro, f_bytes, bar);
                if (has_ga) tma_load_1d(st + (SOFT ? kTRow + kFRow : 0), prm.g_alpha + ro, f_bytes, bar);
            }
        }
    };
    for (int q = 0; q < kRing - 1; ++q) issue(q);

    float a_n[VPT], ax_n = 0.f, gd_n = 0.f;
    int L_n = 0, xcol_n = -1;
    auto fetch = [&](int q) {           // grid values of row q, one row ahead (see the forward kernel)
        const long long row = row0 + q;
        const int n = (int)((unsigned)row / (unsigned)T_len);          // N*T < 2^31 (checked by the launcher)
        L_n = __ldg(prm.lens + n);
        xcol_n = __ldg(prm.xcol + n);
        ax_n = __ldg(prm.a_x + row);
        if (has_gd) gd_n = __ldg(prm.g_delays + row);
        const float* asp = prm.a_sp + (size_t)row * Sp;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            a_n[k] = 0.f;
            if (pat.on(k)) a_n[k] = __ldg(asp + pat.idx(k));
        }
    };
    fetch(0);
    for (int q = 0; q < nrows; ++q) {
        const long long row = row0 + q;
        issue(q + kRing - 1);
        int L = L_n;
        const int xcol = xcol_n;
        const float gd = gd_n;
        float a[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) a[k] = (j0 + k == xcol) ? ax_n : a_n[k];
        const bool own_x = xcol >= j0 && xcol < j0 + VPT;
        if (q + 1 < nrows) fetch(q + 1);
        if (L < 0) L = 0;       // poisoned row: gradients of nothing
        if (any_tma) mbar_wait(&bars[q % kRing], (unsigned)((q / kRing) & 1));
        const unsigned char* st = ring + (q % kRing) * kStage;
        // external dL/d alpha' at this thread's sparse columns
        float gsp[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) gsp[k] = 0.f;
        {
            const float* garow = reinterpret_cast<const float*>(st + (SOFT ? kTRow + kFRow : 0));
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const bool sparse_col = pat.on(k) || (j0 + k == xcol);
                if (sparse_col) gsp[k] = (has_ga ? garow[j0 + k] : 0.f) + gd * (float)(j0 + k + 1);
            }
        }
        if constexpr (SOFT) {
            const T* erow = reinterpret_cast<const T*>(st) + j0;
            float E[VPT];
            load_typed<T, VPT>(erow, E);
#pragma unroll
            for (int k = 0; k < VPT; ++k) E[k] = (j0 + k < L) ? E[k] : -INFINITY;
            float mx = E[0];
#pragma unroll
            for (int k = 1; k < VPT; ++k) mx = fmaxf(mx, E[k]);
            // first thread holding the row maximum (autograd routes max's gradient to the arg-max)
            if (tid == 0) sh_int[q & 1] = 0x7fffffff;
            if constexpr (NW == 1) __syncwarp();
            const float m = block_max<NW>(mx, xs(0), warp, lane);
            float exm[VPT], e[VPT], D[VPT], et = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                exm[k] = (j0 + k < L) ? ex2a((E[k] - m) * kLog2eS) : 0.f;
                e[k] = (j0 + k < L) ? exm[k] + eps : 0.f;
                et += e[k];
                D[k] = et;
            }
            const float2 ep = block_prefix<NW>(et, xs(1), warp, lane);
            if (mx == m && L > 0) atomicMin(&sh_int[q & 1], tid);       // (the reset is a barrier behind)
            float rD[VPT], rr[VPT], R[VPT], rt = 0.f;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k) {
                const bool sparse_col = (pat.on(k) || (j0 + k == xcol)) && (j0 + k < L);
                rD[k] = sparse_col ? fast_rcp(eps + (ep.x + D[k])) : 0.f;
                rr[k] = a[k] * rD[k];
                rt += rr[k];
                R[k] = rt;
            }
            const float2 rp = block_suffix<NW>(rt, xs(2), warp, lane);
            // gb = grad_beta * 1[0 <= b <= 1]; ge1 = gb * R; gR = gb * e; gr = prefix(gR)
            float ge1[VPT], gr[VPT], gt = 0.f;
            {
                const float* gbrow = reinterpret_cast<const float*>(st + kTRow);
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float Rk = rp.x + R[k];
                    const float b = e[k] * Rk;
                    float gb = (has_gb && j0 + k < L) ? gbrow[j0 + k] : 0.f;
                    gb = (b >= 0.0f && b <= 1.0f) ? gb : 0.f;
                    ge1[k] = gb * Rk;
                    gt += gb * e[k];
                    gr[k] = gt;
                }
            }
            const float2 gp = block_prefix<NW>(gt, xs(3), warp, lane);
            // at the sparse columns: d/d alpha' = gr / D ; hD = gr * r / D ( = -gD ); suffix(hD)
            float H[VPT], ht = 0.f;
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k) {
                const float gsoft = (gp.x + gr[k]) * rD[k];
                gsp[k] += gsoft;
                ht += gsoft * rr[k];
                H[k] = ht;
            }
            const float2 hp = block_suffix<NW>(ht, xs(4), warp, lane);
            float gE[VPT], gs = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                gE[k] = (ge1[k] - (hp.x + H[k])) * exm[k];
                gs += gE[k];
            }
            gs = block_sum1<NW>(gs, xs(5), warp, lane);
            if constexpr (NW == 1) __syncwarp();
            if (sh_int[q & 1] == tid) {
                bool done = false;
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (!done && E[k] == m) { gE[k] -= gs; done = true; }
            }
            if (j0 < S) {
                T* out = reinterpret_cast<T*>(prm.g_e) + (size_t)row * S + j0;
                T gv[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) gv[k] = from_f32<T>(j0 + k < L ? gE[k] : 0.f);
                store_typed<T, VPT>(out, S - j0, gv);
            }
            xs.flip();
        }
        // ---- d/d alpha' on the grid
        float* gout = prm.g_sp + (size_t)row * Sp;
#pragma unroll
        for (int k = 0; k < VPT; ++k)
            if (pat.on(k)) gout[pat.idx(k)] = gsp[k];
        if (own_x) {
#pragma unroll
            for (int k = 0; k < VPT; ++k)
                if (j0 + k == xcol) prm.g_x4[row].x = gsp[k];
        }
        if (!SOFT && any_tma) __syncthreads();      // ring slot reuse (the SOFT path has its own barriers)
    }
}


// =============================================================================================
// K2 / K3 for the common geometry: every grid column is the LAST column of some thread
// (ratio a multiple of VPT = 8 and S a multiple of 8 -- ratio 8, what exp/2-mma.sh trains).  The
// grid quantities of a thread are then scalars, which takes the row kernels from ~950 to ~200
// instructions per warp and row (the generic kernels above are issue bound).  The residual of a
// right-padded row on an off-grid column (xcol) is a correction applied by the one thread that
// owns the column.
template <int VPT>
__device__ __forceinline__ float pick(const float (&v)[VPT], int k) {
    float o = v[0];
#pragma unroll
    for (int c = 1; c < VPT; ++c) o = (c == k) ? v[c] : o;
    return o;
}

template <int THREADS, typename T>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 1024 / THREADS : 1))
sparse_row_fwd_last_kernel(const SparseParams prm, int rows_per_cta) {
    constexpr int VPT = 8, NW = THREADS / 32, CAP = THREADS * VPT;
    constexpr int kRing = ring_fwd(CAP);
    constexpr int kRowBytes = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    unsigned char* ring = smem + 128 + 2 * 3 * kXFloats * 4;
    Xs<3> xs{xraw, 0};

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const int rows_total = prm.N * T_len;
    const int row0 = blockIdx.x * rows_per_cta;
    const int nrows = min(rows_per_cta, rows_total - row0);
    if (nrows <= 0) return;
    const int j0 = tid * VPT;
    const bool in_row = j0 < S;                                 // S % 8 == 0: wholly inside or outside
    const int m7 = in_row ? grid_idx(j0 + VPT - 1, Sp, S, r) : -1;       // pooled index of this thread's grid column
    const unsigned row_bytes = (unsigned)(S * sizeof(T));

    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int q) {
        if (warp == 0 && q < nrows) {
            if (elect1()) {
                uint64_t* bar = &bars[q % kRing];
                mbar_expect_tx(bar, row_bytes);
                tma_load_1d(ring + (q % kRing) * kRowBytes, reinterpret_cast<const T*>(prm.e) + (size_t)(row0 + q) * S,
                            row_bytes, bar);
            }
        }
    };
    for (int q = 0; q < kRing - 1; ++q) issue(q);

    int n = row0 / T_len, t_in = row0 - n * T_len;              // (n, t) of the row being processed
    int L = __ldg(prm.lens + n), xcol = __ldg(prm.xcol + n);
    float a7_n = m7 >= 0 ? __ldg(prm.a_sp + (size_t)row0 * Sp + m7) : 0.f;
    float ax_n = __ldg(prm.a_x + row0);
    const float qnan = __int_as_float(0x7fc00000);
    const T* pp_base = reinterpret_cast<const T*>(prm.pp);

    for (int q = 0; q < nrows; ++q) {
        const int row = row0 + q;
        issue(q + kRing - 1);
        float a7 = a7_n;
        const float ax = ax_n;
        if (q + 1 < nrows) {                                     // next row's grid values, one row ahead
            if (m7 >= 0) a7_n = __ldg(prm.a_sp + (size_t)(row + 1) * Sp + m7);
            ax_n = __ldg(prm.a_x + row + 1);
        }
        const int nl = min(max(L - j0, 0), VPT);                 // live columns of this thread
        const int kx = (xcol >= j0 && xcol < j0 + VPT) ? xcol - j0 : -1;
        // warp-uniform: every lane wholly live, no off-grid residual column, no poisoned row
        const bool all_full = __all_sync(kFull, nl == VPT && kx < 0) && L >= 0;
        if (L < 0) a7 = qnan;                                    // poisoned row
        if (in_row && prm.alpha != nullptr) {
            float* arow = prm.alpha + (size_t)row * S + j0;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = make_float4(0.f, 0.f, 0.f, a7);
            if (L < 0) { lo = make_float4(qnan, qnan, qnan, qnan); hi = lo; }
            if (kx >= 0) {
                float v[VPT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, a7};
#pragma unroll
                for (int k = 0; k < VPT; ++k) v[k] = (k == kx) ? ax : v[k];
                lo = make_float4(v[0], v[1], v[2], v[3]); hi = make_float4(v[4], v[5], v[6], v[7]);
            }
            *reinterpret_cast<float4*>(arow) = lo;
            *reinterpret_cast<float4*>(arow + 4) = hi;
        }
        if (in_row) {
            if (prm.p_dense != nullptr) {
                T pv[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) pv[k] = from_f32<T>(0.f);
                if (m7 >= 0) pv[VPT - 1] = __ldg(pp_base + (size_t)row * Sp + m7);
                store_typed<T, VPT>(reinterpret_cast<T*>(prm.p_dense) + (size_t)row * S + j0, VPT, pv);
            }
        }
        mbar_wait(&bars[q % kRing], (unsigned)((q / kRing) & 1));
        float E[VPT];
        load_typed<T, VPT>(reinterpret_cast<const T*>(ring + (q % kRing) * kRowBytes) + j0, E);
        if (!all_full) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) E[k] = (k < nl) ? E[k] : -INFINITY;
        }
        float mx = fmaxf(fmaxf(fmaxf(E[0], E[1]), fmaxf(E[2], E[3])), fmaxf(fmaxf(E[4], E[5]), fmaxf(E[6], E[7])));
        float m = block_max<NW>(mx, xs(0), warp, lane);
        if (L <= 0) m = 0.f;                                    // no live column: exp(-inf - 0) = 0 everywhere
        const float nm = -m * kLog2eS;
        float e[VPT], D[VPT], et = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            // padded columns (E = -inf) get e = eps like the reference's masked_fill(-1e8) path; they sit
            // right of every live column, so no live prefix sees them, and their beta is 0 (R = 0 there)
            e[k] = ex2a(__fmaf_rn(E[k], kLog2eS, nm)) + eps;
            et += e[k];
            D[k] = et;
        }
        if (nl == 0) et = 0.f;
        const float2 ep = block_prefix<NW>(et, xs(1), warp, lane);
        // r = alpha / (eps + cumsum(e)) at the grid column (and the residual column); R = suffix(r)
        float r7 = (nl == VPT) ? a7 * fast_rcp(eps + (ep.x + D[VPT - 1])) : 0.f;
        float rx = 0.f;
        if (!all_full && kx >= 0) rx = ax * fast_rcp(eps + (ep.x + pick<VPT>(D, kx)));
        const float2 rp = block_suffix<NW>(r7 + rx, xs(2), warp, lane);
        if (in_row) {
            const float Rt = rp.x + r7;
            float b[VPT];
            if (all_full) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) b[k] = fminf(fmaxf(e[k] * Rt, 0.0f), 1.0f);
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float Rk = (kx >= 0 && k <= kx) ? Rt + rx : Rt;
                    b[k] = fminf(fmaxf(e[k] * Rk, 0.0f), 1.0f);
                    b[k] = (k < nl) ? b[k] : 0.f;
                    if (L < 0) b[k] = qnan;
                }
            }
            float* brow = prm.beta + (size_t)row * S + j0;
            *reinterpret_cast<float4*>(brow) = make_float4(b[0], b[1], b[2], b[3]);
            *reinterpret_cast<float4*>(brow + 4) = make_float4(b[4], b[5], b[6], b[7]);
        }
        xs.flip();
        if (++t_in == T_len) {                                  // next (batch, head) row: new geometry
            t_in = 0; ++n;
            if (q + 1 < nrows) { L = __ldg(prm.lens + n); xcol = __ldg(prm.xcol + n); }
        }
    }
}

#ifndef SIMULST_K3_THREADS_PER_SM
#define SIMULST_K3_THREADS_PER_SM 1024
#endif
template <int THREADS, typename T>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? SIMULST_K3_THREADS_PER_SM / THREADS : 1))
sparse_row_bwd_last_kernel(const SparseParams prm, int rows_per_cta) {
    constexpr int VPT = 8, NW = THREADS / 32, CAP = THREADS * VPT;
    constexpr int kRing = ring_bwd(CAP);
    constexpr int kTRow = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    constexpr int kFRow = CAP * 4;
    constexpr int kStage = kTRow + 2 * kFRow;                           // energy, grad_beta, grad_alpha
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);
    int* sh_int = reinterpret_cast<int*>(smem + 128 + 2 * 6 * kXFloats * 4);
    unsigned char* ring = smem + 128 + 2 * 6 * kXFloats * 4 + 256;
    Xs<6> xs{xraw, 0};

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const int rows_total = prm.N * T_len;
    const int row0 = blockIdx.x * rows_per_cta;
    const int nrows = min(rows_per_cta, rows_total - row0);
    if (nrows <= 0) return;
    const int j0 = tid * VPT;
    const bool in_row = j0 < S;
    const int m7 = in_row ? grid_idx(j0 + VPT - 1, Sp, S, r) : -1;
    const unsigned t_bytes = (unsigned)(S * sizeof(T)), f_bytes = (unsigned)(S * 4);
    const bool has_ga = prm.g_alpha != nullptr;
    const bool has_gb = prm.g_beta != nullptr;
    const bool has_gd = prm.g_delays != nullptr;

    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int q) {
        if (warp == 0 && q < nrows) {
            if (elect1()) {
                uint64_t* bar = &bars[q % kRing];
                unsigned char* st = ring + (q % kRing) * kStage;
                const size_t ro = (size_t)(row0 + q) * S;
                mbar_expect_tx(bar, t_bytes + (has_gb ? f_bytes : 0u) + (has_ga ? f_bytes : 0u));
                tma_load_1d(st, reinterpret_cast<const T*>(prm.e) + ro, t_bytes, bar);
                if (has_gb) tma_load_1d(st + kTRow, prm.g_beta + ro, f_bytes, bar);
                if (has_ga) tma_load_1d(st + kTRow + kFRow, prm.g_alpha + ro, f_bytes, bar);
            }
        }
    };
    for (int q = 0; q < kRing - 1; ++q) issue(q);

    int n = row0 / T_len, t_in = row0 - n * T_len;
    int L = __ldg(prm.lens + n), xcol = __ldg(prm.xcol + n);
    float a7_n = m7 >= 0 ? __ldg(prm.a_sp + (size_t)row0 * Sp + m7) : 0.f;
    float ax_n = __ldg(prm.a_x + row0);
    float gd_n = has_gd ? __ldg(prm.g_delays + row0) : 0.f;

    for (int q = 0; q < nrows; ++q) {
        const int row = row0 + q;
        issue(q + kRing - 1);
        const float a7 = a7_n, ax = ax_n, gd = gd_n;
        if (q + 1 < nrows) {
            if (m7 >= 0) a7_n = __ldg(prm.a_sp + (size_t)(row + 1) * Sp + m7);
            ax_n = __ldg(prm.a_x + row + 1);
            if (has_gd) gd_n = __ldg(prm.g_delays + row + 1);
        }
        const int Lc = max(L, 0);                                // poisoned row: gradients of nothing
        const int nl = min(max(Lc - j0, 0), VPT);
        const int kx = (xcol >= j0 && xcol < j0 + VPT) ? xcol - j0 : -1;
        // warp-uniform: every lane wholly live and no off-grid residual column in this warp -- the
        // element-wise loops then need no per-column selects (the common case: 28 % fewer instructions)
        const bool all_full = __all_sync(kFull, nl == VPT && kx < 0);
        if (tid == 0) sh_int[q & 1] = 0x7fffffff;
        if constexpr (NW == 1) __syncwarp();
        mbar_wait(&bars[q % kRing], (unsigned)((q / kRing) & 1));
        const unsigned char* st = ring + (q % kRing) * kStage;
        float E[VPT];
        load_typed<T, VPT>(reinterpret_cast<const T*>(st) + j0, E);
        if (!all_full) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) E[k] = (k < nl) ? E[k] : -INFINITY;
        }
        const float mx = fmaxf(fmaxf(fmaxf(E[0], E[1]), fmaxf(E[2], E[3])), fmaxf(fmaxf(E[4], E[5]), fmaxf(E[6], E[7])));
        float m = block_max<NW>(mx, xs(0), warp, lane);
        if (L <= 0) m = 0.f;                                    // no live column: exp(-inf - 0) = 0 everywhere
        const float nm = -m * kLog2eS;
        float exm[VPT], D[VPT], et = 0.f;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            exm[k] = ex2a(__fmaf_rn(E[k], kLog2eS, nm));
            et += exm[k] + eps;
            D[k] = et;
        }
        if (nl == 0) et = 0.f;
        const float2 ep = block_prefix<NW>(et, xs(1), warp, lane);
        if (mx == m && nl > 0) atomicMin(&sh_int[q & 1], tid);
        const float rD7 = (nl == VPT) ? fast_rcp(eps + (ep.x + D[VPT - 1])) : 0.f;
        const float r7 = a7 * rD7;
        float rDx = 0.f, rx = 0.f;
        if (!all_full && kx >= 0) { rDx = fast_rcp(eps + (ep.x + pick<VPT>(D, kx))); rx = ax * rDx; }
        const float2 rp = block_suffix<NW>(r7 + rx, xs(2), warp, lane);
        // gb = grad_beta * 1[0 <= b <= 1] ; ge1 = gb * R ; gR = gb * e ; gr = prefix(gR)
        const float Rt = rp.x + r7;
        float ge1[VPT], gr[VPT], gt = 0.f;
        {
            float gbv[VPT];
            if (has_gb) {
                const float4* gbrow = reinterpret_cast<const float4*>(st + kTRow) + 2 * tid;
                const float4 g0 = gbrow[0], g1 = gbrow[1];
                gbv[0] = g0.x; gbv[1] = g0.y; gbv[2] = g0.z; gbv[3] = g0.w;
                gbv[4] = g1.x; gbv[5] = g1.y; gbv[6] = g1.z; gbv[7] = g1.w;
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) gbv[k] = 0.f;
            }
            if (all_full) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float ek = exm[k] + eps;
                    const float b = ek * Rt;
                    const float gb = (b >= 0.0f && b <= 1.0f) ? gbv[k] : 0.f;
                    ge1[k] = gb * Rt;
                    gt += gb * ek;
                    gr[k] = gt;
                }
            } else {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const float Rk = (kx >= 0 && k <= kx) ? Rt + rx : Rt;
                    const float ek = exm[k] + eps;
                    const float b = ek * Rk;
                    float gb = (k < nl) ? gbv[k] : 0.f;
                    gb = (b >= 0.0f && b <= 1.0f) ? gb : 0.f;
                    ge1[k] = gb * Rk;
                    gt += gb * ek;
                    gr[k] = gt;
                }
            }
        }
        const float2 gp = block_prefix<NW>(gt, xs(3), warp, lane);
        // grid column: d/d alpha' = gr / D ; hD = that * r ; H = suffix(hD)
        const float gsoft7 = (gp.x + gr[VPT - 1]) * rD7;
        float gsoftx = 0.f;
        if (!all_full && kx >= 0) gsoftx = (gp.x + pick<VPT>(gr, kx)) * rDx;
        const float hx = gsoftx * rx;
        const float h7 = gsoft7 * r7;
        const float2 hp = block_suffix<NW>(h7 + hx, xs(4), warp, lane);
        const float Ht = hp.x + h7;
        float gE[VPT], gs = 0.f;
        if (all_full) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                gE[k] = (ge1[k] - Ht) * exm[k];
                gs += gE[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const float Hk = (kx >= 0 && k <= kx) ? Ht + hx : Ht;
                gE[k] = (ge1[k] - Hk) * exm[k];
                gs += gE[k];
            }
        }
        gs = block_sum1<NW>(gs, xs(5), warp, lane);
        if constexpr (NW == 1) __syncwarp();
        if (sh_int[q & 1] == tid) {
            bool done = false;
#pragma unroll
            for (int k = 0; k < VPT; ++k)
                if (!done && E[k] == m) { gE[k] -= gs; done = true; }
        }
        if (in_row) {
            T gv[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) gv[k] = from_f32<T>(gE[k]);          // exm = 0 at padded columns
            store_typed<T, VPT>(reinterpret_cast<T*>(prm.g_e) + (size_t)row * S + j0, VPT, gv);
            // d/d alpha' on the grid: soft-attention term + external grad_alpha + delays term
            if (m7 >= 0) {
                const float ga = has_ga ? reinterpret_cast<const float*>(st + kTRow + kFRow)[j0 + VPT - 1] : 0.f;
                prm.g_sp[(size_t)row * Sp + m7] = gsoft7 + ga + gd * (float)(j0 + VPT);
            }
            if (kx >= 0) {
                const float ga = has_ga ? reinterpret_cast<const float*>(st + kTRow + kFRow)[xcol] : 0.f;
                prm.g_x4[row].x = gsoftx + ga + gd * (float)(xcol + 1);
            }
        }
        xs.flip();
        if (++t_in == T_len) {
            t_in = 0; ++n;
            if (q + 1 < nrows) { L = __ldg(prm.lens + n); xcol = __ldg(prm.xcol + n); }
        }
    }
}

// =============================================================================================
// K4: expected alignment on the grid, backward (A.2 + A.4).  One CTA per (batch, head) row, steps
// walked in reverse, the recurrence gradient ("carry") in registers.
template <int NW, int EPT, typename T>
__global__ void __launch_bounds__(NW * 32) sparse_alpha_bwd_kernel(const SparseParams prm) {
    constexpr int THREADS = NW * 32;
    __shared__ float xraw[2 * 5 * kXFloats];
    __shared__ float sh_b[2];
    Xs<5> xs{xraw, 0};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool replace = prm.mask == nullptr;
    const int m0 = tid * EPT;
    T* gout = reinterpret_cast<T*>(prm.g_pp) + (size_t)n * T_len * Sp;

    int L = prm.lens[n];
    const int xcol = prm.xcol[n];
    if (L < 0) {        // poisoned row: zero gradients
        for (size_t q = tid; q < (size_t)T_len * Sp; q += THREADS) gout[q] = from_f32<T>(0.f);
        return;
    }
    int mp_m = -1;
    if (mp) {
        if (replace) mp_m = Sp - 1;
        else if (L > 0 && xcol < 0) mp_m = grid_idx(L - 1, Sp, S, r);
    }
    const bool has_mp = mp && (mp_m >= 0 || xcol >= 0);

    float W[EPT];
    bool valid[EPT], live[EPT];
    const double log1e = log((double)(1.0f + eps));
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int m = m0 + k;
        valid[k] = m < Sp;
        const int col = valid[k] ? grid_col(m, Sp, S, r) : 0;
        live[k] = valid[k] && col < L;
        W[k] = valid[k] ? (float)exp((double)(1 + col - m) * log1e) : 1.0f;
    }
    const int k_mp = (mp_m >= m0 && mp_m < m0 + EPT) ? mp_m - m0 : -1;

    const T* gpp = reinterpret_cast<const T*>(prm.pp) + (size_t)n * T_len * Sp;
    const float* asp = prm.a_sp + (size_t)n * T_len * Sp;
    const float* gsp = prm.g_sp + (size_t)n * T_len * Sp;
    const float4* gx = prm.g_x4 + (size_t)n * T_len;
    const float* side = prm.side != nullptr ? prm.side + (size_t)n * T_len * 2 : nullptr;

    float carry[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) carry[k] = 0.f;

    // operands of step i are fetched one iteration ahead
    unsigned raw_p[EPT];
    float G_n[EPT], am1_n[EPT];
    float side_sum_n = 0.f, side_raw_prev_n = 0.f, gx_n = 0.f;
    auto fetch = [&](int i) {
        const T* prow = gpp + (size_t)i * Sp + m0;
        const float* grow = gsp + (size_t)i * Sp + m0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            raw_p[k] = valid[k] ? ldg_raw<T>(prow + k) : 0u;
            G_n[k] = valid[k] ? __ldg(grow + k) : 0.f;
            am1_n[k] = (valid[k] && i > 0) ? __ldg(asp + (size_t)(i - 1) * Sp + m0 + k) : 0.f;
        }
        if (has_mp) {
            side_sum_n = __ldg(side + 2 * i + 1);
            if (i > 0) side_raw_prev_n = __ldg(side + 2 * (i - 1));
            if (xcol >= 0) gx_n = __ldg(&gx[i].x);
        }
    };
    if (T_len > 0) fetch(T_len - 1);

    for (int i = T_len - 1; i >= 0; --i) {
        float p[EPT], G[EPT], am1[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const float v = raw_to_f32<T>(raw_p[k]);
            p[k] = live[k] ? v : 0.f;
            G[k] = live[k] ? G_n[k] : 0.f;
            am1[k] = am1_n[k];
        }
        const float side_sum = side_sum_n, side_raw_prev = side_raw_prev_n, gx_i = gx_n;
        if (i > 0) fetch(i - 1);
        // alpha_{i-1} as the recurrence saw it: undo mass preservation on the stored row
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) am1[k] = (S == 1 && m0 + k == 0) ? 1.0f : 0.0f;
        } else if (k_mp >= 0) {
#pragma unroll
            for (int k = 0; k < EPT; ++k)
                if (k == k_mp) am1[k] = side_raw_prev;
        }
        // ---- mass preservation backward: ga_j = g'_j - ok * g'_mp (REPLACE: 0 at the column itself)
        float gmp = 0.f;
        if (has_mp) {
            if (xcol >= 0) {
                gmp = gx_i;
            } else {
                if (k_mp >= 0) {
#pragma unroll
                    for (int k = 0; k < EPT; ++k)
                        if (k == k_mp) sh_b[i & 1] = G[k];
                }
                if constexpr (NW == 1) __syncwarp(); else __syncthreads();
                gmp = sh_b[i & 1];
            }
            const float ok = (side_sum >= 0.0f && side_sum <= 1.0f) ? 1.0f : 0.0f;
            gmp *= ok;
        }
        // ---- recompute the forward quantities of step i
        float x[EPT], xe[EPT], xt = 1.0f;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            x[k] = valid[k] ? ((1.0f - p[k]) + eps) : 1.0f;
            xe[k] = xt;
            xt *= x[k];
        }
        const float xoff = block_prefix_mul<NW>(xt, xs(0), warp, lane);
        float cp[EPT], rc[EPT], pass[EPT], P[EPT], u[EPT], sl[EPT], ut = 0.f;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            cp[k] = W[k] * (xoff * xe[k]);
            const float c = fminf(fmaxf(cp[k], eps), 1.0f);
            rc[k] = fast_rcp(c);
            pass[k] = (c == cp[k]) ? 1.0f : 0.0f;           // 1[eps <= cp <= 1]
            P[k] = p[k] * cp[k];
            u[k] = am1[k] * rc[k];
            ut += u[k];
            sl[k] = ut;
        }
        const float2 up = block_prefix<NW>(ut, xs(1), warp, lane);
        const float s_off = up.x + ((i == 0 && S != 1) ? 1.0f : 0.0f);
        // ---- g = ga + carry ; gz = g * 1[0 <= z <= 1] ; gP = gz * s ; gs = gz * P ; gu = suffix(gs)
        float gP[EPT], gsl[EPT], gst = 0.f;
#pragma unroll
        for (int k = EPT - 1; k >= 0; --k) {
            const float s = s_off + sl[k];
            const float z = P[k] * s;
            float ga = G[k] - gmp;
            if (replace && k == k_mp) ga = 0.f;
            const float g = live[k] ? ga + carry[k] : 0.f;
            const float gz = (z >= 0.0f && z <= 1.0f) ? g : 0.f;
            gP[k] = gz * s;
            gst += gz * P[k];
            gsl[k] = gst;
        }
        const float2 gup = block_suffix<NW>(gst, xs(2), warp, lane);
        // ---- carry = gu / c ; gc = -carry * u ; gcp = gP*p + gc*pass ; gA = gcp * cp ; gL = excl suffix(gA)
        float gAl[EPT], gAt = 0.f;
#pragma unroll
        for (int k = EPT - 1; k >= 0; --k) {
            const float gu = gup.x + gsl[k];
            carry[k] = gu * rc[k];
            const float gcp = gP[k] * p[k] - (carry[k] * u[k]) * pass[k];
            gAl[k] = gAt;                                   // exclusive local suffix
            gAt += gcp * cp[k];
        }
        const float2 gAp = block_suffix<NW>(gAt, xs(3), warp, lane);
        T* orow = gout + (size_t)i * Sp + m0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const float gL = gAp.x + gAl[k];
            const float o = gP[k] * cp[k] - gL * fast_rcp(x[k]);
            if (valid[k]) orow[k] = from_f32<T>(live[k] ? o : 0.f);
        }
        xs.flip();
    }
}


// =============================================================================================
// K1 / K4, warp-specialised (rows whose grid fits one warp: Sp <= 32 * EPT).  A single warp that
// does everything issues ~240 (forward) / ~500 (backward) instructions per step in order, at one
// instruction per ~4.4 cycles: the recurrence waits behind work that does not depend on it.  Here
// the step is split by DEPENDENCE, one role per warp group, connected by shared-memory rings with
// full / empty mbarriers:
//   producers  (NP warps, step i -> warp i % NP)   everything that does not depend on the recurrence:
//              loads, cumprod scan, clamp, 1/c, P (backward: also u, s, clamp masks, upstream gradient)
//   chain      (warp 0)   the recurrence itself: one multiply, one scan, one clamp per step
//   consumers  (NC warps) everything downstream: row sums, mass preservation, delays, stores
//              (backward: the exclusive suffix scan and the gradient of the pooled p_choose)
// The chain warp's step is ~45 instructions around one shuffle scan.
constexpr int kWsDepth = 8;

template <int EPT, int NP, int NC, typename T>
__global__ void __launch_bounds__((1 + NP + NC) * 32) sparse_alpha_fwd_ws_kernel(const SparseParams prm) {
    constexpr int CAPP = 32 * EPT, THREADS = (1 + NP + NC) * 32;
    __shared__ __align__(16) float ringP[kWsDepth][CAPP], ringRc[kWsDepth][CAPP], ringA[kWsDepth][CAPP];
    __shared__ __align__(8) uint64_t fullA[kWsDepth], fullC[kWsDepth], emptyC[kWsDepth];
    __shared__ int sh_int[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool replace = prm.mask == nullptr;

    const RowGeom geo = row_geometry<THREADS>(prm, n, sh_int);
    if (tid == 0) { prm.lens[n] = geo.L; prm.xcol[n] = geo.xcol; }
    float* asp = prm.a_sp + (size_t)n * T_len * Sp;
    float* ax = prm.a_x + (size_t)n * T_len;
    float4* info = prm.mp_info + (size_t)n * T_len;
    if (geo.L < 0) {
        if (tid == 0 && prm.status != nullptr) atomicOr(prm.status, SIMULST_ST_NOT_RIGHT_PADDED);
        const float qnan = __int_as_float(0x7fc00000);
        for (size_t q = tid; q < (size_t)T_len * Sp; q += THREADS) asp[q] = qnan;
        for (int q = tid; q < T_len; q += THREADS) { ax[q] = qnan; info[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
        return;
    }
    if (tid == 0) {
        for (int s2 = 0; s2 < kWsDepth; ++s2) { mbar_init(&fullA[s2], 1); mbar_init(&fullC[s2], 1); mbar_init(&emptyC[s2], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    const int L = geo.L;
    const int m0 = lane * EPT;
    bool valid[EPT], live[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        valid[k] = m0 + k < Sp;
        live[k] = valid[k] && grid_col(m0 + k, Sp, S, r) < L;
    }

    if (warp == 0) {
        // ------------------------------------------------ chain: u = alpha_{i-1} / c ; s = prefix(u) ; alpha_i = clamp(P s)
        float a_prev[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) a_prev[k] = (S == 1 && m0 + k == 0) ? 1.0f : 0.0f;
        for (int i = 0; i < T_len; ++i) {
            const int slot = i % kWsDepth;
            mbar_wait(&fullA[slot], (unsigned)((i / kWsDepth) & 1));
            float P[EPT], rc[EPT], u[EPT], ut = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < EPT / 4; ++c4) {
                const float4 pv = *reinterpret_cast<const float4*>(&ringP[slot][m0 + 4 * c4]);
                const float4 rv = *reinterpret_cast<const float4*>(&ringRc[slot][m0 + 4 * c4]);
                P[4 * c4] = pv.x; P[4 * c4 + 1] = pv.y; P[4 * c4 + 2] = pv.z; P[4 * c4 + 3] = pv.w;
                rc[4 * c4] = rv.x; rc[4 * c4 + 1] = rv.y; rc[4 * c4 + 2] = rv.z; rc[4 * c4 + 3] = rv.w;
            }
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                ut += a_prev[k] * rc[k];
                u[k] = ut;
            }
            const float inc = wscan_prefix_add(ut);
            const float s_off = wprev(inc, 0.f) + ((i == 0 && S != 1) ? 1.0f : 0.0f);
            float a[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                a[k] = fminf(fmaxf(P[k] * (s_off + u[k]), 0.0f), 1.0f);
                a_prev[k] = a[k];
            }
#pragma unroll
            for (int c4 = 0; c4 < EPT / 4; ++c4)
                *reinterpret_cast<float4*>(&ringA[slot][m0 + 4 * c4]) = make_float4(a[4 * c4], a[4 * c4 + 1], a[4 * c4 + 2], a[4 * c4 + 3]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullC[slot]);
        }
    } else if (warp <= NP) {
        // ------------------------------------------------ producers: P = p * cp, 1 / clamp(cp)
        const int w = warp - 1;
        float W[EPT];
        const double log1e = log((double)(1.0f + eps));
#pragma unroll
        for (int k = 0; k < EPT; ++k)
            W[k] = valid[k] ? (float)exp((double)(1 + grid_col(m0 + k, Sp, S, r) - (m0 + k)) * log1e) : 1.0f;
        const T* gpp = reinterpret_cast<const T*>(prm.pp) + (size_t)n * T_len * Sp;
        unsigned umax = 0u;
        unsigned raw[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) raw[k] = (valid[k] && w < T_len) ? ldg_raw<T>(gpp + (size_t)w * Sp + m0 + k) : 0u;
        for (int i = w; i < T_len; i += NP) {
            float p[EPT], xe[EPT], xt = 1.0f;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                umax = max(umax, sizeof(T) == 4 ? raw[k] : raw[k] << 16);
                const float v = raw_to_f32<T>(raw[k]);
                p[k] = live[k] ? v : 0.f;
                xe[k] = xt;
                xt *= (1.0f - p[k]) + eps;
            }
            if (i + NP < T_len) {
#pragma unroll
                for (int k = 0; k < EPT; ++k) raw[k] = valid[k] ? ldg_raw<T>(gpp + (size_t)(i + NP) * Sp + m0 + k) : 0u;
            }
            const float xoff = wprev(wscan_prefix_mul(xt), 1.0f);
            float P[EPT], rc[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const float cp = W[k] * (xoff * xe[k]);
                P[k] = p[k] * cp;
                rc[k] = fast_rcp(fminf(fmaxf(cp, eps), 1.0f));
            }
            const int slot = i % kWsDepth;
            mbar_wait_relaxed(&emptyC[slot], (unsigned)(((i / kWsDepth) & 1) ^ 1));
#pragma unroll
            for (int c4 = 0; c4 < EPT / 4; ++c4) {
                *reinterpret_cast<float4*>(&ringP[slot][m0 + 4 * c4]) = make_float4(P[4 * c4], P[4 * c4 + 1], P[4 * c4 + 2], P[4 * c4 + 3]);
                *reinterpret_cast<float4*>(&ringRc[slot][m0 + 4 * c4]) = make_float4(rc[4 * c4], rc[4 * c4 + 1], rc[4 * c4 + 2], rc[4 * c4 + 3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[slot]);
        }
        // prob_check / safe_cumprod's sign check: bit patterns above 1.0f are NaN, > 1 or negative; the
        // exact classification only runs when the cheap test trips
        if (prm.status != nullptr && __any_sync(kFull, umax > 0x3f800000u)) {
            unsigned bits = 0u;
            for (int i = w; i < T_len; i += NP)
                for (int m = lane; m < Sp; m += 32) {
                    const float v = to_f32<T>(gpp[(size_t)i * Sp + m]);
                    bits |= prob_bits(v) | ((((1.0f - v) + eps) < 0.f) ? SIMULST_ST_NEGPROD : 0u);
                }
            bits = __reduce_or_sync(kFull, bits);
            if (lane == 0 && bits) atomicOr(prm.status, bits);
        }
    } else {
        // ------------------------------------------------ consumers: row sums, mass preservation, delays, stores
        const int w = warp - 1 - NP;
        const bool has_res = mp && (geo.mp_m >= 0 || geo.xcol >= 0);
        const int k_mp = (mp && geo.mp_m >= m0 && geo.mp_m < m0 + EPT) ? geo.mp_m - m0 : -1;
        float wcolm[EPT], summ[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const bool excl = replace && k == k_mp;         // REPLACE: the residual excludes the column itself
            summ[k] = excl ? 0.f : 1.0f;
            wcolm[k] = excl ? 0.f : (valid[k] ? (float)(grid_col(m0 + k, Sp, S, r) + 1) : 0.f);
        }
        const float w_mp = geo.mp_m >= 0 ? (float)(grid_col(geo.mp_m, Sp, S, r) + 1) : (float)(geo.xcol + 1);
        const bool vec_store = (Sp % EPT) == 0;
        const bool info_lane = geo.mp_m >= 0 ? (k_mp >= 0) : (lane == 0);      // the lane holding the raw column value
        for (int i = w; i < T_len; i += NC) {
            const int slot = i % kWsDepth;
            mbar_wait_relaxed(&fullC[slot], (unsigned)((i / kWsDepth) & 1));
            float a[EPT];
#pragma unroll
            for (int c4 = 0; c4 < EPT / 4; ++c4) {
                const float4 av = *reinterpret_cast<const float4*>(&ringA[slot][m0 + 4 * c4]);
                a[4 * c4] = av.x; a[4 * c4 + 1] = av.y; a[4 * c4 + 2] = av.z; a[4 * c4 + 3] = av.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyC[slot]);
            float pa = 0.f, pw = 0.f;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                pa += a[k] * summ[k];
                pw += a[k] * wcolm[k];
            }
            pa = warp_sum(pa);
            pw = warp_sum(pw);
            const float res = has_res ? 1.0f - fminf(fmaxf(pa, 0.0f), 1.0f) : 0.f;
            float raw_v = 0.f;
            if (k_mp >= 0) {
#pragma unroll
                for (int k = 0; k < EPT; ++k)
                    if (k == k_mp) { raw_v = a[k]; a[k] = replace ? res : a[k] + res; }
            }
            if (info_lane) {
                info[i] = make_float4(pa, raw_v, 0.f, 0.f);
                if (mp && prm.side != nullptr) {
                    prm.side[((size_t)n * T_len + i) * 2] = raw_v;
                    prm.side[((size_t)n * T_len + i) * 2 + 1] = pa;
                }
            }
            if (lane == 0) {
                ax[i] = (mp && geo.xcol >= 0) ? res : 0.f;
                if (prm.delays != nullptr) prm.delays[(size_t)n * T_len + i] = has_res ? pw + w_mp * res : pw;
            }
            float* arow = asp + (size_t)i * Sp + m0;
            if (vec_store) {
                if (valid[0]) {
#pragma unroll
                    for (int c4 = 0; c4 < EPT / 4; ++c4)
                        *reinterpret_cast<float4*>(arow + 4 * c4) = make_float4(a[4 * c4], a[4 * c4 + 1], a[4 * c4 + 2], a[4 * c4 + 3]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < EPT; ++k)
                    if (valid[k]) arow[k] = a[k];
            }
        }
    }
}

template <int EPT, int NP, int NC, typename T>
__global__ void __launch_bounds__((1 + NP + NC) * 32) sparse_alpha_bwd_ws_kernel(const SparseParams prm) {
    constexpr int CAPP = 32 * EPT, THREADS = (1 + NP + NC) * 32;
    // per slot: producer -> chain {ga, mz*P, 1/c}; chain -> consumer {g, carry}; producer -> consumer {mz*s, p, u*pass, cp, 1/x}
    extern __shared__ __align__(128) unsigned char smem_ws[];
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_ws);
    uint64_t* fullC = fullA + kWsDepth;
    uint64_t* emptyC = fullC + kWsDepth;
    float* ring = reinterpret_cast<float*>(smem_ws + 256);
    auto slot_arr = [&](int slot, int which) { return ring + ((size_t)slot * 10 + which) * CAPP; };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    const int S = prm.S, Sp = prm.Sp, r = prm.r, T_len = prm.T;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool replace = prm.mask == nullptr;
    const int m0 = lane * EPT;
    T* gout = reinterpret_cast<T*>(prm.g_pp) + (size_t)n * T_len * Sp;

    const int L = prm.lens[n];
    const int xcol = prm.xcol[n];
    if (L < 0) {
        for (size_t q = tid; q < (size_t)T_len * Sp; q += THREADS) gout[q] = from_f32<T>(0.f);
        return;
    }
    if (tid == 0) {
        for (int s2 = 0; s2 < kWsDepth; ++s2) { mbar_init(&fullA[s2], 1); mbar_init(&fullC[s2], 1); mbar_init(&emptyC[s2], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    bool valid[EPT], live[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        valid[k] = m0 + k < Sp;
        live[k] = valid[k] && grid_col(m0 + k, Sp, S, r) < L;
    }
    auto ld4 = [&](const float* src, float (&v)[EPT]) {
#pragma unroll
        for (int c4 = 0; c4 < EPT / 4; ++c4) {
            const float4 t4 = *reinterpret_cast<const float4*>(src + m0 + 4 * c4);
            v[4 * c4] = t4.x; v[4 * c4 + 1] = t4.y; v[4 * c4 + 2] = t4.z; v[4 * c4 + 3] = t4.w;
        }
    };
    auto st4 = [&](float* dst, const float (&v)[EPT]) {
#pragma unroll
        for (int c4 = 0; c4 < EPT / 4; ++c4)
            *reinterpret_cast<float4*>(dst + m0 + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
    };
    // steps are walked T-1 .. 0; q = T-1-i is the processing index that selects ring slot and phase

    if (warp == 0) {
        // ------------------------------------------------ chain: g = ga + carry ; gu = suffix(g * mz * P) ; carry = gu / c
        float carry[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) carry[k] = 0.f;
        for (int q = 0; q < T_len; ++q) {
            const int slot = q % kWsDepth;
            mbar_wait(&fullA[slot], (unsigned)((q / kWsDepth) & 1));
            float ga[EPT], mzP[EPT], rc[EPT], g[EPT], gsl[EPT], gst = 0.f;
            ld4(slot_arr(slot, 0), ga);
            ld4(slot_arr(slot, 1), mzP);
            ld4(slot_arr(slot, 2), rc);
#pragma unroll
            for (int k = EPT - 1; k >= 0; --k) {
                g[k] = ga[k] + carry[k];
                gst += g[k] * mzP[k];
                gsl[k] = gst;
            }
            const float off = wnext(wscan_suffix_add(gst), 0.f);
#pragma unroll
            for (int k = 0; k < EPT; ++k) carry[k] = (off + gsl[k]) * rc[k];
            // (the consumer released this slot D steps ago -- implied by fullA, waited for directly so that
            //  every write-after-read on the ring has its own barrier edge)
            mbar_wait(&emptyC[slot], (unsigned)(((q / kWsDepth) & 1) ^ 1));
            st4(slot_arr(slot, 3), g);
            st4(slot_arr(slot, 4), carry);
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullC[slot]);
        }
    } else if (warp <= NP) {
        // ------------------------------------------------ producers: recompute the forward quantities of a step
        const int w = warp - 1;
        int mp_m = -1;
        if (mp) {
            if (replace) mp_m = Sp - 1;
            else if (L > 0 && xcol < 0) mp_m = grid_idx(L - 1, Sp, S, r);
        }
        const bool has_mp = mp && (mp_m >= 0 || xcol >= 0);
        const int k_mp = (mp_m >= m0 && mp_m < m0 + EPT) ? mp_m - m0 : -1;
        const int mp_lane = mp_m >= 0 ? mp_m / EPT : 0;
        float W[EPT];
        const double log1e = log((double)(1.0f + eps));
#pragma unroll
        for (int k = 0; k < EPT; ++k)
            W[k] = valid[k] ? (float)exp((double)(1 + grid_col(m0 + k, Sp, S, r) - (m0 + k)) * log1e) : 1.0f;
        const T* gpp = reinterpret_cast<const T*>(prm.pp) + (size_t)n * T_len * Sp;
        const float* asp = prm.a_sp + (size_t)n * T_len * Sp;
        const float* gsp = prm.g_sp + (size_t)n * T_len * Sp;
        const float4* info = prm.mp_info + (size_t)n * T_len;
        const float4* gx4 = prm.g_x4 + (size_t)n * T_len;
        // operands of this warp's next step, fetched one own-step ahead
        unsigned raw_p[EPT];
        float am1_n[EPT], G_n[EPT], ssum_n = 0.f, rawprev_n = 0.f, gx_n = 0.f;
        auto fetch = [&](int i) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                raw_p[k] = valid[k] ? ldg_raw<T>(gpp + (size_t)i * Sp + m0 + k) : 0u;
                G_n[k] = valid[k] ? __ldg(gsp + (size_t)i * Sp + m0 + k) : 0.f;
                am1_n[k] = (valid[k] && i > 0) ? __ldg(asp + (size_t)(i - 1) * Sp + m0 + k) : 0.f;
            }
            if (has_mp) {
                ssum_n = __ldg(&info[i].x);
                if (i > 0 && mp_m >= 0) rawprev_n = __ldg(&info[i - 1].y);
                if (xcol >= 0) gx_n = __ldg(&gx4[i].x);
            }
        };
        if (w < T_len) fetch(T_len - 1 - w);
        for (int q = w; q < T_len; q += NP) {
            const int i = T_len - 1 - q;
            float p[EPT], G[EPT], am1[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                p[k] = live[k] ? raw_to_f32<T>(raw_p[k]) : 0.f;
                G[k] = live[k] ? G_n[k] : 0.f;
                am1[k] = am1_n[k];
            }
            const float ssum = ssum_n, rawprev = rawprev_n, gx_i = gx_n;
            if (q + NP < T_len) fetch(i - NP);
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < EPT; ++k) am1[k] = (S == 1 && m0 + k == 0) ? 1.0f : 0.0f;
            } else if (k_mp >= 0) {                       // undo mass preservation on the stored row
#pragma unroll
                for (int k = 0; k < EPT; ++k) am1[k] = (k == k_mp) ? rawprev : am1[k];
            }
            float gmp = 0.f;
            if (has_mp) {
                float cand = gx_i;
                if (xcol < 0) {
                    float own = 0.f;
#pragma unroll
                    for (int k = 0; k < EPT; ++k) own = (k == k_mp) ? G[k] : own;
                    cand = __shfl_sync(kFull, own, mp_lane);
                }
                gmp = (ssum >= 0.0f && ssum <= 1.0f) ? cand : 0.f;
            }
            float x[EPT], xe[EPT], xt = 1.0f;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                x[k] = (1.0f - p[k]) + eps;
                xe[k] = xt;
                xt *= x[k];
            }
            const float xoff = wprev(wscan_prefix_mul(xt), 1.0f);
            float cp[EPT], rc[EPT], pass[EPT], P[EPT], u[EPT], sl[EPT], ut = 0.f;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                cp[k] = W[k] * (xoff * xe[k]);
                const float cc = fminf(fmaxf(cp[k], eps), 1.0f);
                rc[k] = fast_rcp(cc);
                pass[k] = (cc == cp[k]) ? 1.0f : 0.0f;
                P[k] = p[k] * cp[k];
                u[k] = am1[k] * rc[k];
                ut += u[k];
                sl[k] = ut;
            }
            const float s_off = wprev(wscan_prefix_add(ut), 0.f) + ((i == 0 && S != 1) ? 1.0f : 0.0f);
            float ga[EPT], mzP[EPT], ms[EPT], up[EPT], rx[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const float s = s_off + sl[k];
                const float z = P[k] * s;
                const float mz = (z >= 0.0f && z <= 1.0f) ? 1.0f : 0.0f;
                float gak = G[k] - gmp;
                if (replace && k == k_mp) gak = 0.f;
                ga[k] = live[k] ? gak : 0.f;
                mzP[k] = mz * P[k];
                ms[k] = mz * s;
                up[k] = u[k] * pass[k];
                rx[k] = fast_rcp(x[k]);
            }
            const int slot = q % kWsDepth;
            mbar_wait_relaxed(&emptyC[slot], (unsigned)(((q / kWsDepth) & 1) ^ 1));
            if (q >= kWsDepth) mbar_wait_relaxed(&fullC[slot], (unsigned)(((q / kWsDepth) - 1) & 1));
            st4(slot_arr(slot, 0), ga);
            st4(slot_arr(slot, 1), mzP);
            st4(slot_arr(slot, 2), rc);
            st4(slot_arr(slot, 5), ms);
            st4(slot_arr(slot, 6), p);
            st4(slot_arr(slot, 7), up);
            st4(slot_arr(slot, 8), cp);
            st4(slot_arr(slot, 9), rx);
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA[slot]);
        }
    } else {
        // ------------------------------------------------ consumers: gradient of the pooled p_choose of a step
        const int w = warp - 1 - NP;
        const bool vec_store = (Sp % EPT) == 0 && sizeof(T) * EPT >= 8;
        for (int q = w; q < T_len; q += NC) {
            const int i = T_len - 1 - q;
            const int slot = q % kWsDepth;
            mbar_wait_relaxed(&fullA[slot], (unsigned)((q / kWsDepth) & 1));
            mbar_wait_relaxed(&fullC[slot], (unsigned)((q / kWsDepth) & 1));
            float g[EPT], carry[EPT], ms[EPT], p[EPT], up[EPT], cp[EPT], rx[EPT];
            ld4(slot_arr(slot, 3), g);
            ld4(slot_arr(slot, 4), carry);
            ld4(slot_arr(slot, 5), ms);
            ld4(slot_arr(slot, 6), p);
            ld4(slot_arr(slot, 7), up);
            ld4(slot_arr(slot, 8), cp);
            ld4(slot_arr(slot, 9), rx);
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyC[slot]);
            // gP = gz * s ; gcp = gP * p - carry * u * 1[eps <= cp <= 1] ; gA = gcp * cp ; gL = exclusive suffix(gA)
            float gP[EPT], gAl[EPT], gAt = 0.f;
#pragma unroll
            for (int k = EPT - 1; k >= 0; --k) {
                gP[k] = g[k] * ms[k];
                const float gcp = gP[k] * p[k] - carry[k] * up[k];
                gAl[k] = gAt;
                gAt += gcp * cp[k];
            }
            const float off = wnext(wscan_suffix_add(gAt), 0.f);
            T* orow = gout + (size_t)i * Sp + m0;
            __align__(16) T ov[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const float o = gP[k] * cp[k] - (off + gAl[k]) * rx[k];
                ov[k] = from_f32<T>(live[k] ? o : 0.f);
            }
            if (vec_store) {
                if (valid[0]) {
                    if constexpr (sizeof(T) * EPT == 8) {
                        *reinterpret_cast<uint2*>(orow) = *reinterpret_cast<const uint2*>(ov);
                    } else {
#pragma unroll
                        for (int c4 = 0; c4 < (int)(sizeof(T) * EPT / 16); ++c4)
                            reinterpret_cast<uint4*>(orow)[c4] = reinterpret_cast<const uint4*>(ov)[c4];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < EPT; ++k)
                    if (valid[k]) orow[k] = ov[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------ launchers
template <typename T>
int launch_alpha(const SparseParams& prm, bool bwd, cudaStream_t st) {
#define SIMULST_SP_ALPHA(NWv, EP)                                                              \
    {                                                                                          \
        if (bwd) sparse_alpha_bwd_kernel<NWv, EP, T><<<prm.N, NWv * 32, 0, st>>>(prm);         \
        else sparse_alpha_fwd_kernel<NWv, EP, T><<<prm.N, NWv * 32, 0, st>>>(prm);             \
        return check_launch();                                                                 \
    }
    const int Sp = prm.Sp;
    // rows whose grid fits one warp: the warp-specialised kernels (2 producers + chain + 2 consumers forward,
    // 3 + chain + 2 backward)
    if (Sp <= 256 && g_sparse_variant == 0) {
        static bool attr_done[2][64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const int which = Sp <= 128 ? 0 : 1;
        if (!bwd) {
            if (which == 0) sparse_alpha_fwd_ws_kernel<4, 2, 2, T><<<prm.N, 5 * 32, 0, st>>>(prm);
            else sparse_alpha_fwd_ws_kernel<8, 2, 2, T><<<prm.N, 5 * 32, 0, st>>>(prm);
            return check_launch();
        }
        // (2 + 2 / 3 + 2 helper warps: more of them only take issue slots from the chain warps --
        // measured 3+3: 41 us, 4+4: 41 us forward; 4+3: 85 us, 6+4: 78 us backward)
        const size_t smem = 256 + (size_t)kWsDepth * 10 * (which == 0 ? 128 : 256) * 4;
        auto k4 = sparse_alpha_bwd_ws_kernel<4, 3, 2, T>;
        auto k8 = sparse_alpha_bwd_ws_kernel<8, 3, 2, T>;
        if (!attr_done[which][dev & 63]) {
            if (cudaFuncSetAttribute(which == 0 ? k4 : k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return SIMULST_E_LAUNCH;
            }
            attr_done[which][dev & 63] = true;
        }
        if (which == 0) k4<<<prm.N, 6 * 32, smem, st>>>(prm);
        else k8<<<prm.N, 6 * 32, smem, st>>>(prm);
        return check_launch();
    }
    if (Sp <= 128) SIMULST_SP_ALPHA(1, 4)
    if (Sp <= 256) SIMULST_SP_ALPHA(1, 8)
    if (Sp <= 512) SIMULST_SP_ALPHA(2, 8)
    if (Sp <= 1024) SIMULST_SP_ALPHA(4, 8)
    if (Sp <= 2048) SIMULST_SP_ALPHA(8, 8)
    SIMULST_SP_ALPHA(16, 8)         // Sp <= 4096
#undef SIMULST_SP_ALPHA
}

inline int rows_per_cta(long long rows) {
    // enough CTAs for ~8 per SM, at most 16 rows each so the TMA ring has something to run ahead on
    const long long want = rows / (148 * 8);
    return (int)std::max<long long>(1, std::min<long long>(16, want));
}

template <int THREADS, int VPT, typename T, bool SOFT>
int launch_row_impl(const SparseParams& prm, bool bwd, cudaStream_t st) {
    constexpr int CAP = THREADS * VPT;
    constexpr int kTRow = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    const long long rows = (long long)prm.N * prm.T;
    if (rows >= (1ll << 31)) return SIMULST_E_SHAPE;
    const int rpc = rows_per_cta(rows);
    const unsigned grid = (unsigned)((rows + rpc - 1) / rpc);
    static bool attr_done[4][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if constexpr (SOFT && VPT == 8) {
        // every grid column is the last column of a thread: the lean kernels
        if (prm.r % VPT == 0 && prm.S % VPT == 0) {
            if (!bwd) {
                const size_t smem = 128 + 2 * 3 * kXFloats * 4 + (size_t)ring_fwd(CAP) * kTRow;
                auto kern = sparse_row_fwd_last_kernel<THREADS, T>;
                if (!attr_done[2][dev & 63]) {
                    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                        cudaGetLastError();
                        return SIMULST_E_LAUNCH;
                    }
                    attr_done[2][dev & 63] = true;
                }
                kern<<<grid, THREADS, smem, st>>>(prm, rpc);
            } else {
                const size_t smem = 128 + 2 * 6 * kXFloats * 4 + 256 + (size_t)ring_bwd(CAP) * (kTRow + 2 * CAP * 4);
                auto kern = sparse_row_bwd_last_kernel<THREADS, T>;
                if (!attr_done[3][dev & 63]) {
                    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                        cudaGetLastError();
                        return SIMULST_E_LAUNCH;
                    }
                    attr_done[3][dev & 63] = true;
                }
                kern<<<grid, THREADS, smem, st>>>(prm, rpc);
            }
            return check_launch();
        }
    }
    if (!bwd) {
        const size_t smem = 128 + 2 * 3 * kXFloats * 4 + (SOFT ? (size_t)ring_fwd(CAP) * kTRow : 0);
        auto kern = sparse_row_fwd_kernel<THREADS, VPT, T, SOFT>;
        if (!attr_done[0][dev & 63]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return SIMULST_E_LAUNCH;
            }
            attr_done[0][dev & 63] = true;
        }
        kern<<<grid, THREADS, smem, st>>>(prm, rpc);
    } else {
        const size_t stage = (SOFT ? (size_t)kTRow + CAP * 4 : 0) + (size_t)CAP * 4;
        const size_t smem = 128 + 2 * 6 * kXFloats * 4 + 256 + (size_t)ring_bwd(CAP) * stage;
        auto kern = sparse_row_bwd_kernel<THREADS, VPT, T, SOFT>;
        if (!attr_done[1][dev & 63]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return SIMULST_E_LAUNCH;
            }
            attr_done[1][dev & 63] = true;
        }
        kern<<<grid, THREADS, smem, st>>>(prm, rpc);
    }
    return check_launch();
}

template <typename T, bool SOFT>
int launch_row(const SparseParams& prm, bool bwd, cudaStream_t st) {
    const int S = prm.S;
    if (S <= 256) return launch_row_impl<32, 8, T, SOFT>(prm, bwd, st);
    if (S <= 512) return launch_row_impl<64, 8, T, SOFT>(prm, bwd, st);
    if (S <= 1024) return launch_row_impl<128, 8, T, SOFT>(prm, bwd, st);
    if (S <= 2048) return launch_row_impl<256, 8, T, SOFT>(prm, bwd, st);
    if (S <= 4096) return launch_row_impl<512, 8, T, SOFT>(prm, bwd, st);
    return SIMULST_E_SHAPE;
}

template <typename T>
int run_sparse(const SparseParams& prm, bool bwd, cudaStream_t st) {
    const bool soft = (prm.flags & SIMULST_MMA_SOFT) != 0u;
    if (!bwd) {
        int rc = launch_alpha<T>(prm, false, st);
        if (rc != SIMULST_OK) return rc;
        return soft ? launch_row<T, true>(prm, false, st) : launch_row<T, false>(prm, false, st);
    }
    int rc = soft ? launch_row<T, true>(prm, true, st) : launch_row<T, false>(prm, true, st);
    if (rc != SIMULST_OK) return rc;
    return launch_alpha<T>(prm, true, st);
}

}  // namespace

int mma_sparse_run(const SparseParams& prm, int dtype, bool bwd, cudaStream_t st) {
    switch (dtype) {
        case SIMULST_F32: return run_sparse<float>(prm, bwd, st);
        case SIMULST_BF16: return run_sparse<__nv_bfloat16>(prm, bwd, st);
        default: return run_sparse<__half>(prm, bwd, st);
    }
}

}  // namespace simulst
