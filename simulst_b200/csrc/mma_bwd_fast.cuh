// MMA training backward, dense fast path: same mathematics as mma_bwd.cuh (see its header for
// the derivation) in five block-wide exchanges per target step (the
// product scan and the e-scan share one: the row max is reduced one step ahead), for rows whose
// threads are wholly inside or wholly outside the row -- S a multiple of the per-thread element
// count, up to 16 warps x 12 elements (S <= 6144) -- with no padding mask or a right-padding
// mask (template modes RAGGED / MASKED below), rows staged by TMA, hard or infinite-lookback soft
// attention.  Everything else (other masks, chunkwise windows, unaligned rows) stays with the
// generic kernel.  Written for instruction count, because that kernel is issue bound:
//   * element-wise arithmetic on float2 pairs (FADD2 / FMUL2 / FFMA2: one issue slot per two
//     source positions); only the thread-local scan chains stay scalar;
//   * warp scans use the shuffle's own predicate (no lane compares, no selects), and the scans
//     that share a barrier advance level by level in one asm block so their latencies overlap;
//   * cross-warp offsets are formed with 0/1 float weights (FMUL + FFMA) instead of predicated
//     adds, and clamp masks are 0/1 floats folded into the packed multiplies: the loop keeps no
//     long-lived predicates (the generic kernel spends ~140 instructions per step on LOP3 /
//     ISETP traffic spilling and re-deriving predicates);
//   * the mass-preservation correction ok*g'_last is formed by every thread from the block
//     totals of the e- and gR-scans (g'_last = gA_last + sum(gR)/(eps + sum(e))), so the
//     recurrence needs ONE suffix scan (the generic kernel splits the term off by linearity and
//     pays a second scan);
//   * the arg-max search runs only in the thread that holds the row maximum.
// Every floating-point operation that feeds a clamp mask (cp, s, z) is performed in the same
// order as in the forward kernels, so the recomputed masks agree with the forward pass; without
// mass preservation the results are bit-identical to mma_bwd_kernel.
#pragma once

#include "mma_bwd.cuh"
#include "mma_scan.cuh"

namespace simulst {

constexpr int kFastMaxWarps = 16;          // CTA size limit of the fast path
constexpr int kFastMaxWeightWarps = 8;    // up to here cross-warp offsets use 0/1 register weights

// ---- fused warp-scan levels (shuffle predicate = "source lane exists")
// x: prefix product, m: max over the warp
template <int D>
__device__ __forceinline__ void lvl_mul_max(float& x, float& m) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.bfly.b32 t1, %1, %2, 31, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "max.f32 %1, %1, t1;\n\t}"
        : "+f"(x), "+f"(m)
        : "n"(D));
}
// x: prefix product, e: prefix sum (one shared shuffle predicate)
template <int D>
__device__ __forceinline__ void lvl_mul_add(float& x, float& e) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t1, %1, %2, 0, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "@q0 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(x), "+f"(e)
        : "n"(D));
}
// a: suffix sum (distance DA), s: butterfly sum, m: butterfly max (distance DS)
template <int DA, int DS>
__device__ __forceinline__ void lvl_dn_sum_max(float& a, float& s, float& m) {
    asm volatile("{\n\t.reg .f32 t0, t1, t2;\n\t.reg .pred q0;\n\t"
        "shfl.sync.down.b32 t0|q0, %0, %3, 31, 0xffffffff;\n\t"
        "shfl.sync.bfly.b32 t1, %1, %4, 31, 0xffffffff;\n\t"
        "shfl.sync.bfly.b32 t2, %2, %4, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "add.rn.f32 %1, %1, t1;\n\t"
        "max.f32 %2, %2, t2;\n\t}"
        : "+f"(a), "+f"(s), "+f"(m)
        : "n"(DA), "n"(DS));
}
// a: prefix sum, b: suffix sum
template <int D>
__device__ __forceinline__ void lvl_up_dn(float& a, float& b) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t1|q1, %1, %2, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "@q1 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(a), "+f"(b)
        : "n"(D));
}
// a, b: suffix sums
template <int D>
__device__ __forceinline__ void lvl_dn_dn(float& a, float& b) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0;\n\t"
        "shfl.sync.down.b32 t0|q0, %0, %2, 31, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t1, %1, %2, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "@q0 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(a), "+f"(b)
        : "n"(D));
}
// a: suffix sum (distance DA), s: butterfly sum (distance DS)
template <int DA, int DS>
__device__ __forceinline__ void lvl_dn_sum(float& a, float& s) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0;\n\t"
        "shfl.sync.down.b32 t0|q0, %0, %2, 31, 0xffffffff;\n\t"
        "shfl.sync.bfly.b32 t1, %1, %3, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(a), "+f"(s)
        : "n"(DA), "n"(DS));
}
// value of the previous / next lane, `ident` at the warp edge
__device__ __forceinline__ float nb_prev(float v, float ident) {
    float o;
    asm volatile("{\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 %0|q, %1, 1, 0, 0xffffffff;\n\t"
        "@!q mov.f32 %0, %2;\n\t}"
        : "=&f"(o)
        : "f"(v), "f"(ident));
    return o;
}
__device__ __forceinline__ float nb_next(float v, float ident) {
    float o;
    asm volatile("{\n\t.reg .pred q;\n\t"
        "shfl.sync.down.b32 %0|q, %1, 1, 31, 0xffffffff;\n\t"
        "@!q mov.f32 %0, %2;\n\t}"
        : "=&f"(o)
        : "f"(v), "f"(ident));
    return o;
}

// global load whose result the compiler treats as thread-varying (keeps a prefetched scalar in
// a vector register instead of converting it to a uniform register -- and waiting -- at once)
__device__ __forceinline__ float ldg_opaque(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- cross-warp combination with 0/1 float weights
template <int NW>
struct WarpWeights {
    static constexpr int kN = (NW > 1 && NW <= kFastMaxWeightWarps) ? NW - 1 : 1;
    float lt[kN];   // lt[w]   = 1 if w < warp      (w = 0 .. NW-2)
    float gt[kN];   // gt[w-1] = 1 if w > warp      (w = 1 .. NW-1)
    __device__ __forceinline__ explicit WarpWeights(int warp) {
        if constexpr (NW > kFastMaxWeightWarps) { lt[0] = 0.f; gt[0] = 0.f; return; }
#pragma unroll
        for (int w = 0; w + 1 < NW && w < kN; ++w) {
            lt[w] = (w < warp) ? 1.0f : 0.0f;
            gt[w] = (w + 1 > warp) ? 1.0f : 0.0f;
        }
        if (NW == 1) { lt[0] = 0.f; gt[0] = 0.f; }
    }
};
template <int NW>
__device__ __forceinline__ void load_totals(const float* __restrict__ wt, float (&t)[NW]) {
    if constexpr (NW == 1) {
        t[0] = wt[0];
    } else if constexpr (NW == 2) {
        const float2 v = *reinterpret_cast<const float2*>(wt);
        t[0] = v.x; t[1] = v.y;
    } else {
        // (a slot holds kXStride floats, so the last vector may reach past NW; the extra lanes are dropped)
#pragma unroll
        for (int q = 0; q < (NW + 3) / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(wt + 4 * q);
            if (4 * q < NW) t[4 * q] = v.x;
            if (4 * q + 1 < NW) t[4 * q + 1] = v.y;
            if (4 * q + 2 < NW) t[4 * q + 2] = v.z;
            if (4 * q + 3 < NW) t[4 * q + 3] = v.w;
        }
    }
}
// sum of the warp totals strictly before this warp (same association as combine_prefix)
template <int NW>
__device__ __forceinline__ float off_prefix(const float (&t)[NW], const WarpWeights<NW>& ww) {
    if constexpr (NW == 1) return 0.f;
    float acc = t[0] * ww.lt[0];
#pragma unroll
    for (int w = 1; w + 1 < NW; ++w) acc = __fmaf_rn(t[w], ww.lt[w], acc);
    return acc;
}
template <int NW>
__device__ __forceinline__ float sum_all(const float (&t)[NW]) {
    float acc = t[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) acc += t[w];
    return acc;
}
// sum of the warp totals strictly after this warp (same association as combine_suffix)
template <int NW>
__device__ __forceinline__ float off_suffix(const float (&t)[NW], const WarpWeights<NW>& ww) {
    if constexpr (NW == 1) return 0.f;
    float acc = t[NW - 1] * ww.gt[NW - 2];
#pragma unroll
    for (int w = NW - 2; w >= 1; --w) acc = __fmaf_rn(t[w], ww.gt[w - 1], acc);
    return acc;
}
// product of the warp totals strictly before this warp
template <int NW>
__device__ __forceinline__ float off_prefix_mul(const float (&t)[NW], const WarpWeights<NW>& ww) {
    if constexpr (NW == 1) return 1.0f;
    float acc = __fmaf_rn(t[0], ww.lt[0], 1.0f - ww.lt[0]);
#pragma unroll
    for (int w = 1; w + 1 < NW; ++w) acc *= __fmaf_rn(t[w], ww.lt[w], 1.0f - ww.lt[w]);
    return acc;
}

// ---- block-level combination of the per-warp totals of one exchange slot.  Up to 8 warps:
// every thread reads all totals (one or two LDS.128) and weights them; more warps: lane w of
// every warp takes total w and the warp scans them with shuffles (the association of
// combine_prefix / combine_suffix in common.cuh, so results match the generic kernels).
template <int NW>
__device__ __forceinline__ float xw_off_prefix(const float* __restrict__ wt, const WarpWeights<NW>& ww, int warp,
                                               int lane, float* total) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        float t[NW];
        load_totals<NW>(wt, t);
        if (total != nullptr) *total = sum_all<NW>(t);
        return off_prefix<NW>(t, ww);
    } else {
        const float v = (lane < NW) ? wt[lane] : 0.f;
        const float inc = wscan_prefix_add(v);
        if (total != nullptr) *total = __shfl_sync(kFull, inc, NW - 1);
        return __shfl_sync(kFull, inc - v, warp);
    }
}
template <int NW>
__device__ __forceinline__ float xw_off_suffix(const float* __restrict__ wt, const WarpWeights<NW>& ww, int warp, int lane) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        float t[NW];
        load_totals<NW>(wt, t);
        return off_suffix<NW>(t, ww);
    } else {
        const float v = (lane < NW) ? wt[lane] : 0.f;
        const float inc = wscan_suffix_add(v);
        return __shfl_sync(kFull, inc - v, warp);
    }
}
template <int NW>
__device__ __forceinline__ float xw_off_prefix_mul(const float* __restrict__ wt, const WarpWeights<NW>& ww, int warp, int lane) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        float t[NW];
        load_totals<NW>(wt, t);
        return off_prefix_mul<NW>(t, ww);
    } else {
        const float v = (lane < NW) ? wt[lane] : 1.0f;
        const float inc = wscan_prefix_mul(v);
        const float exc = nb_prev(inc, 1.0f);
        return __shfl_sync(kFull, exc, warp);
    }
}
template <int NW>
__device__ __forceinline__ float fxw_sum(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        float t[NW];
        load_totals<NW>(wt, t);
        return sum_all<NW>(t);
    } else {
        return warp_sum((lane < NW) ? wt[lane] : 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float fxw_max(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        float t[NW];
        load_totals<NW>(wt, t);
        float m = t[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, t[w]);
        return m;
    } else {
        return warp_max((lane < NW) ? wt[lane] : -INFINITY);
    }
}
// sum of all totals, read by a single thread (no warp collectives)
template <int NW>
__device__ __forceinline__ float xw_sum_one(const float* __restrict__ wt) {
    float acc = wt[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) acc += wt[w];
    return acc;
}
template <int NW>
__device__ __forceinline__ int xw_min_int(const int* __restrict__ wt, int lane) {
    if constexpr (NW <= kFastMaxWeightWarps) {
        int m = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = min(m, wt[w]);
        return m;
    } else {
        return __reduce_min_sync(kFull, (lane < NW) ? wt[lane] : 0x7fffffff);
    }
}

// 1[0 <= v <= 1] as a 0/1 float weight, exactly: v*(1-v) >= 0.  (1-v is exact next to 1, the
// product cannot change sign by rounding, -0 passes like the reference's v >= 0, NaN fails.)
// One packed FMA + one packed multiply per pair, then one FSET per element.
__device__ __forceinline__ float2 unit_mask2(float2 v) {
    const float2 w = mul2(v, fma2(v, f2(-1.0f), f2(1.0f)));
    return make_float2(w.x >= 0.0f ? 1.0f : 0.0f, w.y >= 0.0f ? 1.0f : 0.0f);
}

// VPT consecutive elements of a staged shared-memory row as float2 pairs
template <typename T, int VPT>
__device__ __forceinline__ void lds_row2(const void* __restrict__ row, int j0, float2 (&out)[VPT / 2]) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value && VPT % 8 == 0) {
        // bf16 -> fp32 is a 16-bit shift: one SHL / one LOP3 per element
#pragma unroll
        for (int q = 0; q < VPT / 8; ++q) {
            const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(row) + j0 + 8 * q);
            out[4 * q + 0] = make_float2(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u));
            out[4 * q + 1] = make_float2(__uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
            out[4 * q + 2] = make_float2(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xffff0000u));
            out[4 * q + 3] = make_float2(__uint_as_float(w.w << 16), __uint_as_float(w.w & 0xffff0000u));
        }
    } else {
        float tmp[VPT];
        lds_row<T, VPT>(reinterpret_cast<const T*>(row), j0, tmp);
#pragma unroll
        for (int q = 0; q < VPT / 2; ++q) out[q] = make_float2(tmp[2 * q], tmp[2 * q + 1]);
    }
}

// max over VPT consecutive elements of a staged row; 16-bit rows are reduced with packed
// 16-bit max instructions and converted once
template <typename T, int VPT>
__device__ __forceinline__ float lds_row_max(const void* __restrict__ row, int j0) {
    if constexpr (sizeof(T) == 2 && VPT % 8 == 0) {
        using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(row) + j0);
        T2 acc;
#pragma unroll
        for (int q = 0; q < VPT / 8; ++q) {
            const uint4 w = src[q];
            const T2 a = __hmax2(*reinterpret_cast<const T2*>(&w.x), *reinterpret_cast<const T2*>(&w.y));
            const T2 b = __hmax2(*reinterpret_cast<const T2*>(&w.z), *reinterpret_cast<const T2*>(&w.w));
            const T2 c = __hmax2(a, b);
            acc = q == 0 ? c : __hmax2(acc, c);
        }
        return fmaxf(to_f32<T>(acc.x), to_f32<T>(acc.y));
    } else {
        float2 v[VPT / 2];
        lds_row2<T, VPT>(row, j0, v);
        float m = fmaxf(v[0].x, v[0].y);
#pragma unroll
        for (int q = 1; q < VPT / 2; ++q) m = fmaxf(m, fmaxf(v[q].x, v[q].y));
        return m;
    }
}

// Shared-memory layout (compile-time): header (mbarriers + exchange area), a 3-deep ring of
// stages {p, energy, grad_alpha, grad_beta} and a (stages+1)-deep ring of alpha rows.  The alpha row of
// step i-1 is read by two consecutive iterations (as the recurrence input of step i, then as
// the soft-attention weights alpha'_{i-1} of step i-1), hence the extra slot.
template <int CAP, typename T, bool SOFT>
struct FastLayout {
    static constexpr int kHeader = 128 + 3 * kXSlots * kXStride * 4 + 128;      // mbarriers, 3 exchange buffers
    static constexpr int kTRow = (CAP * (int)sizeof(T) + 127) / 128 * 128;
    static constexpr int kFRow = (CAP * 4 + 127) / 128 * 128;
    static constexpr int kOffP = 0;
    static constexpr int kOffE = kTRow;
    static constexpr int kOffGA = kOffE + (SOFT ? kTRow : 0);
    static constexpr int kOffGB = kOffGA + kFRow;
    static constexpr int kStage = kOffGB + (SOFT ? kFRow : 0);
    // 3 stages (copies issued two steps ahead) when that fits the 227 KB of an SM, else 2
    static constexpr int kStages = (kHeader + 3 * kStage + 4 * kFRow <= 227 * 1024) ? 3 : 2;
    static constexpr int kAlphaSlots = kStages + 1;
    static constexpr int kOffAlpha = kHeader + kStages * kStage;
    static constexpr int kTotal = kOffAlpha + kAlphaSlots * kFRow;
};

// RAGGED: S < THREADS*VPT with S a multiple of VPT (every thread is wholly inside or wholly
// outside the row).  The tails [S, CAP) of all ring slots are initialised once to neutral values
// (p = 0, energy = -inf, alpha = grads = 0) -- the bulk copies only ever write [0, S) -- so the
// step loop needs no per-element bounds checks: outside threads compute on neutral data, their
// e-term is zeroed by one multiply, and they skip the stores.
// DELAYS: the gradient of the expected delays is compiled in (dense rows: separate instantiation;
// ragged rows: always compiled in, run-time flag).
// MASKED (with RAGGED): padding_mask is a RIGHT-padding mask (the caller's promise, flag
// SIMULST_MMA_RIGHT_PADDING): row n is live on [0, L_n).  As in the forward kernel only the live bytes
// of every row are copied, the ring slots behind them are neutral from a one-time initialisation
// (p = 0, energy = -inf, alpha = grads = 0), and the thread that owns column L_n - 1 neutralises the
// copy overhang of the NEXT step's rows before the step's last barrier -- the step loop itself loads
// without per-element tests.  Gradients beyond L_n are stored as zeros, and mass preservation follows
// the reference's right-padding rule (residual ADDED at L_n - 1).
// SHIFT (with RAGGED and MASKED): input rows need not be 16-byte multiples and S need not be a multiple
// of VPT -- every staged row is the 16-byte aligned superset of the real one, read at its byte offset
// (lds_row2_sh); live length min(S, mask length), the mask optional (without one: the no-mask rule,
// column S-1 REPLACED).  grad_p / grad_energy rows must be 16-byte pitched, pitch >= roundup(S, VPT)
// (padding columns receive zeros); the CTA needs 16 spare columns (S + 16 <= THREADS*VPT).
template <int THREADS, int VPT, typename T, bool SOFT, bool RAGGED, bool DELAYS, bool MASKED = false, bool SHIFT = false>
__global__ void __launch_bounds__(THREADS, (THREADS * VPT <= 1024 ? 4 : (THREADS <= 160 ? 3 : (THREADS <= 256 ? 2 : 1))))
mma_bwd_fast_kernel(const MmaParams prm) {
    constexpr int NW = THREADS / kWarp;
    constexpr int H = VPT / 2;
    // exchange buffers per phase -- soft: X1 0, X3 1, X4 0, X5 1, X6 2 ; hard (no X4): X1 0, X3 1, X5 0, X6 1.
    // A buffer is rewritten only after a barrier that follows its last read.
    constexpr int B3 = 1, B4 = 0, B5 = SOFT ? 1 : 0, B6 = SOFT ? 2 : 1;
    using L = FastLayout<THREADS * VPT, T, SOFT>;
    constexpr int NS = L::kStages;
    constexpr int NA = L::kAlphaSlots;
    static_assert(NW <= kFastMaxWarps && VPT % 4 == 0, "fast path: at most 16 warps");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xraw = reinterpret_cast<float*>(smem + 128);     // [3 buffers][kXSlots][kXStride]
    unsigned char* stage0 = smem + L::kHeader;
    unsigned char* alpha0 = smem + L::kOffAlpha;
    auto xs = [&](int buf, int slot) -> float* { return xraw + (buf * kXSlots + slot) * kXStride; };

    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: the compiler then knows it is warp-uniform (uniform registers,
    // uniform branches around the TMA issue code)
    const int warp = __shfl_sync(kFull, tid >> 5, 0);
    const int n = blockIdx.x;
    if (row_filtered_out(prm, n)) return;       // this row belongs to the call's other pass
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const bool has_ga = prm.g_alpha != nullptr;
    const bool has_gb = SOFT && prm.g_beta != nullptr;
    const bool has_gd = DELAYS && prm.g_delays != nullptr;        // gradient of the expected delays: g'_ij += gd_i * (j+1)
    const bool in_row = !RAGGED || j0 < S;              // this thread's VPT columns exist

    // row pitches in elements; the batch stride of a tensor is T * pitch
    // (only the SHIFT instantiation takes pitched tensors; everywhere else the pitch is S, which lets the
    //  compiler share one row offset between all tensors and keep it in the uniform datapath)
    const int ld_p = SHIFT ? prm.ld_p : S, ld_e = SHIFT ? prm.ld_e : S, ld_a = SHIFT ? prm.ld_alpha : S,
              ld_ga = SHIFT ? prm.ld_ga : S, ld_gb = SHIFT ? prm.ld_gb : S, ld_gp = SHIFT ? prm.ld_gp : S,
              ld_ge = SHIFT ? prm.ld_ge : S;
    const size_t nt = (size_t)n * T_len;
    const T* p_in = reinterpret_cast<const T*>(prm.p) + nt * ld_p;
    const T* e_in = SOFT ? reinterpret_cast<const T*>(prm.e) + nt * ld_e : nullptr;
    const float* al_in = prm.alpha + nt * ld_a;
    const float* ga_in = has_ga ? prm.g_alpha + nt * ld_ga : nullptr;
    const float* gb_in = has_gb ? prm.g_beta + nt * ld_gb : nullptr;
    T* gp_out = reinterpret_cast<T*>(prm.g_p) + nt * ld_gp;
    T* ge_out = SOFT ? reinterpret_cast<T*>(prm.g_e) + nt * ld_ge : nullptr;
    // SHIFT: byte offset of a row inside its staged aligned superset, from the low address bits
    const unsigned p_lo = lo32(p_in), e_lo = lo32(e_in), al_lo = lo32(al_in), ga_lo = lo32(ga_in), gb_lo = lo32(gb_in);
    const unsigned p_pitch = (unsigned)ld_p * (unsigned)sizeof(T), e_pitch = (unsigned)ld_e * (unsigned)sizeof(T),
                   al_pitch = (unsigned)ld_a * 4u, ga_pitch = (unsigned)ld_ga * 4u, gb_pitch = (unsigned)ld_gb * 4u;
    (void)p_lo; (void)e_lo; (void)al_lo; (void)ga_lo; (void)gb_lo;
    (void)p_pitch; (void)e_pitch; (void)al_pitch; (void)ga_pitch; (void)gb_pitch;
    // (+ an opaque zero: the compiler must not treat the prefetched side values as uniform, or it
    // converts them to uniform registers -- and waits for the load -- right at the loop top)
    unsigned opaque_zero;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(opaque_zero));
    opaque_zero >>= 5;
    const float* side = mp ? prm.side + (size_t)n * T_len * 2 + opaque_zero : nullptr;

    const WarpWeights<NW> ww(warp);
    constexpr int kIssuers = NW < 5 ? NW : 5;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], kIssuers);
        mbar_fence_init();
    }
    int live_cnt = 0;
    if (MASKED && in_row) {
        if (SHIFT && prm.mask == nullptr) {
            live_cnt = min(VPT, S - j0);
        } else {
            const uint8_t* mrow = prm.mask + (size_t)n * S + j0;
#pragma unroll
            for (int k = 0; k < VPT; ++k)
                if (!SHIFT || j0 + k < S) live_cnt += (mrow[k] == 0) ? 1 : 0;
        }
    }
    if (MASKED) {
        live_cnt = __reduce_add_sync(kFull, live_cnt);
        if (lane == 0) reinterpret_cast<int*>(xs(2, 3))[warp] = live_cnt;
    }
    if (RAGGED && !MASKED && !in_row) {
        // neutral tails (never overwritten by the bulk copies)
        const T ninf = from_f32<T>(-INFINITY), zero = from_f32<T>(0.f);
        for (int st_i = 0; st_i < NS; ++st_i) {
            unsigned char* st = stage0 + st_i * L::kStage;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                reinterpret_cast<T*>(st + L::kOffP)[j0 + k] = zero;
                if (SOFT) reinterpret_cast<T*>(st + L::kOffE)[j0 + k] = ninf;
                reinterpret_cast<float*>(st + L::kOffGA)[j0 + k] = 0.f;
                if (SOFT) reinterpret_cast<float*>(st + L::kOffGB)[j0 + k] = 0.f;
            }
        }
        for (int a_i = 0; a_i < NA; ++a_i)
#pragma unroll
            for (int k = 0; k < VPT; ++k) reinterpret_cast<float*>(alpha0 + a_i * L::kFRow)[j0 + k] = 0.f;
    }
    __syncthreads();
    // live length of the row and this thread's share of it
    int row_len = S;
    if (MASKED) {
        const int* ci = reinterpret_cast<const int*>(xs(2, 3));
        row_len = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) row_len += ci[w];
        // (through a shuffle: the compiler then knows the value is warp-uniform and keeps the byte counts and
        //  addresses derived from it in the uniform datapath)
        row_len = __shfl_sync(kFull, row_len, 0);
    }
    const int nl = MASKED ? max(0, min(VPT, row_len - j0)) : VPT;      // live columns of this thread
    const bool inside = MASKED ? nl > 0 : in_row;
    const int last_col = MASKED ? max(row_len - 1, 0) : S - 1;
    // owner of the mass-preservation column: dense rows REPLACE column S-1 (the last element of
    // its thread); right-padded rows ADD the residual at L-1 (any element of its thread)
    const bool fixer = MASKED && row_len > 0 && last_col >= j0 && last_col < j0 + VPT;     // owner of the last live column
    (void)fixer;
    if constexpr (MASKED) {
        // (a barrier reduction: the compiler then knows the exit is taken by the whole CTA or by nobody)
        if (__syncthreads_or(row_len == 0 ? 1 : 0)) {
            // no live column: every gradient of the row is zero
            float zero_v[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) zero_v[k] = 0.f;
            if (in_row)
                for (int i = 0; i < T_len; ++i) {
                    st_row_t<T, VPT, true>(gp_out + (size_t)i * ld_gp, j0, S, true, zero_v);
                    if (SOFT) st_row_t<T, VPT, true>(ge_out + (size_t)i * ld_ge, j0, S, true, zero_v);
                }
            return;
        }
        if (in_row && nl == 0) {
            // a thread wholly beyond the row: its gradient columns are zero in every step, written here once
            float zero_v[VPT];
#pragma unroll
            for (int k = 0; k < VPT; ++k) zero_v[k] = 0.f;
            for (int i = 0; i < T_len; ++i) {
                st_row_t<T, VPT, true>(gp_out + (size_t)i * ld_gp, j0, S, true, zero_v);
                if (SOFT) st_row_t<T, VPT, true>(ge_out + (size_t)i * ld_ge, j0, S, true, zero_v);
            }
        }
        if (nl < VPT) {
            // neutral ring slots from this thread's columns on (the copies rewrite the live bytes)
            const T ninf = from_f32<T>(-INFINITY), zero = from_f32<T>(0.f);
            for (int st_i = 0; st_i < NS; ++st_i) {
                unsigned char* st = stage0 + st_i * L::kStage;
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    reinterpret_cast<T*>(st + L::kOffP)[j0 + k] = zero;
                    if (SOFT) reinterpret_cast<T*>(st + L::kOffE)[j0 + k] = ninf;
                    reinterpret_cast<float*>(st + L::kOffGA)[j0 + k] = 0.f;
                    if (SOFT) reinterpret_cast<float*>(st + L::kOffGB)[j0 + k] = 0.f;
                }
            }
            for (int a_i = 0; a_i < NA; ++a_i)
#pragma unroll
                for (int k = 0; k < VPT; ++k) reinterpret_cast<float*>(alpha0 + a_i * L::kFRow)[j0 + k] = 0.f;
            fence_proxy_async();
        }
        __syncthreads();
    }
    const bool mp_last = mp && !MASKED && j0 + VPT == S;
    const int k_add = (MASKED && mp && row_len > 0 && last_col >= j0 && last_col < j0 + VPT) ? last_col - j0 : -1;
    // weight of a column in the mass-preservation Jacobian: 0 for the REPLACED column
    const float w_lastcol = mp_last ? 0.0f : 1.0f;
    // SHIFT rows without a mask follow the no-mask rule: the column at k_add is REPLACED, not added to
    const bool mp_repl = SHIFT && prm.mask == nullptr;
    (void)mp_repl;

    const unsigned t_bytes = (unsigned)((MASKED ? row_len : S) * sizeof(T)), f_bytes = (unsigned)((MASKED ? row_len : S) * 4);
    // Producers: iteration q stages step i = T-1-q into stage q % 3 and alpha slot q % 4
    // (alpha_{i-1}); the very first call also brings alpha'_{T-1} into alpha slot 3.  The five
    // bulk copies are dealt round-robin to the warps (copy j -> lane 0 of warp j % NW) so no
    // single warp carries the whole address arithmetic on the step's critical path; every
    // issuing warp arrives on the stage barrier with the byte count of its own copies.
    auto issue = [&](int q) {
        const int i = T_len - 1 - q;
        const int s = q % NS;
        unsigned char* st = stage0 + s * L::kStage;
        uint64_t* bar = &bars[s];
#pragma unroll
        for (int w = 0; w < kIssuers; ++w) {
            if (warp == w) {                    // warp-uniform
                if (elect_one()) {
                    const bool c_p = (0 % NW) == w, c_e = SOFT && (1 % NW) == w, c_a = (2 % NW) == w,
                               c_ga = (3 % NW) == w, c_gb = SOFT && (4 % NW) == w;      // compile-time
                    const bool a_prev = c_a && i > 0, a_first = c_a && SOFT && q == 0;
                    const bool l_ga = c_ga && has_ga, l_gb = c_gb && has_gb;
                    if constexpr (MASKED) {
                        // the live bytes of every row (SHIFT: of its aligned superset); counts differ from row to row
                        unsigned nb[6] = {0u, 0u, 0u, 0u, 0u, 0u};
                        const void* src[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
                        if (c_p) src[0] = tma_span(p_in + (size_t)i * ld_p, t_bytes, nb[0]);
                        if (c_e) src[1] = tma_span(e_in + (size_t)i * ld_e, t_bytes, nb[1]);
                        if (a_prev) src[2] = tma_span(al_in + (size_t)(i - 1) * ld_a, f_bytes, nb[2]);
                        if (a_first) src[3] = tma_span(al_in + (size_t)i * ld_a, f_bytes, nb[3]);
                        if (l_ga) src[4] = tma_span(ga_in + (size_t)i * ld_ga, f_bytes, nb[4]);
                        if (l_gb) src[5] = tma_span(gb_in + (size_t)i * ld_gb, f_bytes, nb[5]);
                        mbar_expect_tx(bar, nb[0] + nb[1] + nb[2] + nb[3] + nb[4] + nb[5]);
                        if (c_p) tma_load_1d(st + L::kOffP, src[0], nb[0], bar);
                        if (c_e) tma_load_1d(st + L::kOffE, src[1], nb[1], bar);
                        if (a_prev) tma_load_1d(alpha0 + (q % NA) * L::kFRow, src[2], nb[2], bar);
                        if (a_first) tma_load_1d(alpha0 + (NA - 1) * L::kFRow, src[3], nb[3], bar);
                        if (l_ga) tma_load_1d(st + L::kOffGA, src[4], nb[4], bar);
                        if (l_gb) tma_load_1d(st + L::kOffGB, src[5], nb[5], bar);
                    } else {
                        const unsigned bytes = (c_p ? t_bytes : 0u) + (c_e ? t_bytes : 0u) + (a_prev ? f_bytes : 0u) +
                                               (a_first ? f_bytes : 0u) + (l_ga ? f_bytes : 0u) + (l_gb ? f_bytes : 0u);
                        mbar_expect_tx(bar, bytes);
                        if (c_p) tma_load_1d(st + L::kOffP, p_in + (size_t)i * ld_p, t_bytes, bar);
                        if (c_e) tma_load_1d(st + L::kOffE, e_in + (size_t)i * ld_e, t_bytes, bar);
                        if (a_prev) tma_load_1d(alpha0 + (q % NA) * L::kFRow, al_in + (size_t)(i - 1) * ld_a, f_bytes, bar);
                        if (a_first) tma_load_1d(alpha0 + (NA - 1) * L::kFRow, al_in + (size_t)i * ld_a, f_bytes, bar);
                        if (l_ga) tma_load_1d(st + L::kOffGA, ga_in + (size_t)i * ld_ga, f_bytes, bar);
                        if (l_gb) tma_load_1d(st + L::kOffGB, gb_in + (size_t)i * ld_gb, f_bytes, bar);
                    }
                }
            }
        }
    };
    for (int q = 0; q < NS - 1 && q < T_len; ++q) issue(q);

    const float one_eps = 1.0f + eps;
    const float2 eps2 = f2(eps);
    const float2 eps2x = f2((MASKED && nl == 0) ? 0.f : eps);      // no eps from the columns of a thread beyond the row
    const float dead_eps = (MASKED && nl > 0 && nl < VPT) ? (float)(VPT - nl) * eps : 0.f;
    (void)dead_eps;
    const float in_f = inside ? 1.0f : 0.0f;
    (void)in_f;
    // masked rows: neutralise the columns >= nl of a freshly loaded pair array
    auto mask_tail = [&](float2 (&v)[H], float fill) {
        if (MASKED && nl < VPT) {
#pragma unroll
            for (int k = 0; k < VPT; ++k)
                if (k >= nl) SIMULST_EL(v, k) = fill;
        }
    };
    // staged row -> float2 pairs.  SHIFT: at the row's byte offset `sh` in the slot (reads clamped to the
    // slot: threads beyond the row read neutral bytes wherever they land)
    auto load_t = [&](const void* row, unsigned sh, float2 (&v)[H]) {
        if constexpr (SHIFT) {
            if (sh == 0u) lds_row2<T, VPT>(row, j0, v);         // CTA-uniform: this row happens to be aligned
            else lds_row2_sh<T, VPT>(row, sh, j0, THREADS * VPT * (int)sizeof(T), v);
        } else {
            lds_row2<T, VPT>(row, j0, v);
        }
    };
    auto load_f = [&](const void* row, unsigned sh, float2 (&v)[H]) {
        if constexpr (SHIFT) {
            if (sh == 0u) lds_row2<float, VPT>(row, j0, v);
            else lds_row2_sh<float, VPT>(row, sh, j0, THREADS * VPT * 4, v);
        } else {
            lds_row2<float, VPT>(row, j0, v);
        }
    };
    // MASKED: the fixer neutralises the copy overhang behind the row's end in a landed row
    auto fix_t = [&](unsigned char* slot, unsigned sh, unsigned pattern) {
        fix_overhang<(int)sizeof(T), SHIFT>(slot, sh + (unsigned)row_len * (unsigned)sizeof(T), pattern,
                                            (unsigned)(THREADS * VPT * sizeof(T)));
    };
    auto fix_f = [&](unsigned char* slot, unsigned sh) {
        fix_overhang<4, SHIFT>(slot, sh + (unsigned)row_len * 4u, 0u, (unsigned)(THREADS * VPT * 4));
    };
    // all rows that arrive with producer iteration q (step i = T-1-q)
    auto fix_stage = [&](int q) {
        const int i = T_len - 1 - q;
        unsigned char* st = stage0 + (q % NS) * L::kStage;
        fix_t(st + L::kOffP, SHIFT ? sh_of(p_lo, i, p_pitch) : 0u, 0u);
        if (SOFT) fix_t(st + L::kOffE, SHIFT ? sh_of(e_lo, i, e_pitch) : 0u, neg_inf_bits<T>());
        if (has_ga) fix_f(st + L::kOffGA, SHIFT ? sh_of(ga_lo, i, ga_pitch) : 0u);
        if (has_gb) fix_f(st + L::kOffGB, SHIFT ? sh_of(gb_lo, i, gb_pitch) : 0u);
        if (i > 0) fix_f(alpha0 + (q % NA) * L::kFRow, SHIFT ? sh_of(al_lo, i - 1, al_pitch) : 0u);
        if (SOFT && q == 0) fix_f(alpha0 + (NA - 1) * L::kFRow, SHIFT ? sh_of(al_lo, i, al_pitch) : 0u);
    };
    // max over this thread's live columns of a staged energy row (row index i_row)
    auto row_max = [&](const void* row, int i_row) -> float {
        float mx;
        if constexpr (!SHIFT) {
            mx = lds_row_max<T, VPT>(row, j0);
        } else {
            float2 v[H];
            load_t(row, sh_of(e_lo, i_row, e_pitch), v);
            mx = fmaxf(v[0].x, v[0].y);
#pragma unroll
            for (int q = 1; q < H; ++q) mx = fmaxf(mx, fmaxf(v[q].x, v[q].y));
        }
        // a thread beyond the row contributes nothing (it may be looking at overhang that is not fixed yet)
        return (MASKED && nl == 0) ? -INFINITY : mx;
    };
    float2 carry[H];
#pragma unroll
    for (int q = 0; q < H; ++q) carry[q] = f2(0.f);

    // mass-preservation side values (row sum of step i, raw last column of step i-1), fetched
    // one iteration ahead so the global-load latency never sits on the step's critical path
    float side_sum = 0.f, side_prev_last = 0.f;
    if (mp) {
        side_sum = side[2 * (T_len - 1) + 1];
        if (T_len > 1) side_prev_last = side[2 * (T_len - 2)];
    }
    const float* gd_row = nullptr;
    float gd_cur = 0.f;
    if constexpr (DELAYS) {
        if (has_gd) {
            gd_row = prm.g_delays + (size_t)n * T_len + opaque_zero;
            gd_cur = gd_row[T_len - 1];
        }
    }

    // row max of the first step's energies (every later one is reduced one iteration ahead)
    float m_cur = 0.f, Emax_cur = -INFINITY;
    if (MASKED && !SOFT) {
        if (fixer) {
            mbar_wait(&bars[0], 0u);
            fix_stage(0);
        }
        __syncthreads();
    }
    if (SOFT) {
        mbar_wait(&bars[0], 0u);
        if (MASKED && fixer) fix_stage(0);
        Emax_cur = row_max(stage0 + L::kOffE, T_len - 1);
        const float wm = warp_max(Emax_cur);
        if (lane == 0) xs(2, 2)[warp] = wm;
        __syncthreads();
        m_cur = fxw_max<NW>(xs(2, 2), lane);
    }

    int s = 0, a_slot = 0;
    unsigned parity = 0u;
    // MASKED: this thread's gradient addresses as running pointers (row i, walking down): with the row index
    // multiplied out at every store the compiler rebuilt the 64-bit row address from %ctaid each step
    T* pgp_run = MASKED ? gp_out + (size_t)(T_len - 1) * ld_gp + j0 : nullptr;
    T* pge_run = (MASKED && SOFT) ? ge_out + (size_t)(T_len - 1) * ld_ge + j0 : nullptr;
    (void)pgp_run; (void)pge_run;
#pragma unroll 1
    for (int qi = 0; qi < T_len; ++qi) {
        const int i = T_len - 1 - qi;
        if (qi + NS - 1 < T_len) issue(qi + NS - 1);
        float gd_next = 0.f;
        if constexpr (DELAYS) {
            if (has_gd && i > 0) gd_next = ldg_opaque(gd_row + i - 1);
        }
        float side_sum_next = 0.f, side_prev_next = 0.f;
        if (mp && i > 0) {
            side_sum_next = ldg_opaque(side + 2 * (i - 1) + 1);
            if (i > 1) side_prev_next = ldg_opaque(side + 2 * (i - 2));
        }
        mbar_wait(&bars[s], parity);
        const unsigned char* st = stage0 + s * L::kStage;
        const unsigned char* a_prev_row = alpha0 + a_slot * L::kFRow;                               // alpha'_{i-1}
        const unsigned char* a_cur_row = alpha0 + (a_slot == 0 ? NA - 1 : a_slot - 1) * L::kFRow;   // alpha'_i
        if (++s == NS) { s = 0; parity ^= 1u; }
        if (++a_slot == NA) a_slot = 0;

        float2 p[H], E[H];
        load_t(st + L::kOffP, sh_of(p_lo, i, p_pitch), p);
        if (SOFT) load_t(st + L::kOffE, sh_of(e_lo, i, e_pitch), E);

        // ================= X1: exclusive cumprod of (1-p)+eps ; D = eps + prefix(e) ; arg-max owner
        // (the row maximum m_cur of this step's energies was reduced during the previous
        // iteration's last exchange, so the product scan and the e-scan share one barrier)
        float2 cp[H];
        float xtot;
        {
            float2 x[H];
#pragma unroll
            for (int q = 0; q < H; ++q) x[q] = add2(fma2(p[q], f2(-1.0f), f2(1.0f)), eps2);
            float run = x[0].x;
            cp[0] = f2(1.0f, run);
#pragma unroll
            for (int q = 1; q < H; ++q) {
                run *= x[q - 1].y; cp[q].x = run;
                run *= x[q].x;     cp[q].y = run;
            }
            xtot = run * x[H - 1].y;
        }
        float2 ex[H], exm[H], rD[H], Dl[H];
        float etot = 0.f;
        const float m = m_cur;
        if (SOFT) {
            const float2 nm = f2(-m), l2e = f2(kLog2e);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float2 tt = mul2(add2(E[q], nm), l2e);
                exm[q] = f2(ex2_approx(tt.x), ex2_approx(tt.y));
                ex[q] = add2(exm[q], eps2x);
                if (RAGGED && !MASKED) ex[q] = mul2(ex[q], f2(in_f));       // no eps from columns beyond the row
            }
#pragma unroll
            for (int q = 0; q < H; ++q) {
                etot += ex[q].x; Dl[q].x = etot;
                etot += ex[q].y; Dl[q].y = etot;
            }
            // the thread the row ends in summed one eps per column beyond it (after its live columns, whose
            // prefixes are untouched); threads wholly beyond the row have eps = 0
            if (MASKED) etot -= dead_eps;
        }
        float xinc = xtot, einc = etot;
        if (SOFT) {
            lvl_mul_add<1>(xinc, einc); lvl_mul_add<2>(xinc, einc); lvl_mul_add<4>(xinc, einc);
            lvl_mul_add<8>(xinc, einc); lvl_mul_add<16>(xinc, einc);
        } else {
            xinc = wscan_prefix_mul(xinc);
        }
        if (SOFT) {
            // first thread holding the row maximum (the element is located at the end of the step)
            const int cand = __reduce_min_sync(kFull, (Emax_cur == m) ? tid : 0x7fffffff);
            if (lane == 31) {
                xs(0, 1)[warp] = einc;
                reinterpret_cast<int*>(xs(0, 2))[warp] = cand;
            }
        }
        if (lane == 31) xs(0, 0)[warp] = xinc;
        const float xexc = nb_prev(xinc, 1.0f);
        float eexc = 0.f;
        if (SOFT) eexc = nb_prev(einc, 0.f);
        __syncthreads();
        float2 rc[H], P[H], pass[H];
        float rD_last = 0.f;
        int amax = 0;
        {
            const float xoff = xw_off_prefix_mul<NW>(xs(0, 0), ww, warp, lane);
            const float2 cbase = f2((one_eps * xoff) * xexc);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                cp[q] = mul2(cbase, cp[q]);
                const float2 c = f2(fminf(fmaxf(cp[q].x, eps), 1.0f), fminf(fmaxf(cp[q].y, eps), 1.0f));
                rc[q] = rcp2(c);
                pass[q] = f2(c.x == cp[q].x ? 1.0f : 0.0f, c.y == cp[q].y ? 1.0f : 0.0f);   // 1[eps <= cp <= 1]
                P[q] = mul2(p[q], cp[q]);
            }
        }
        if (SOFT) {
            amax = xw_min_int<NW>(reinterpret_cast<const int*>(xs(0, 2)), lane);
            float e_all = 0.f;
            const float2 ebase = f2(xw_off_prefix<NW>(xs(0, 1), ww, warp, lane, mp ? &e_all : nullptr) + eexc);
            if (mp) rD_last = fast_rcp(eps + e_all);
#pragma unroll
            for (int q = 0; q < H; ++q) rD[q] = rcp2(add2(eps2, add2(ebase, Dl[q])));
        }

        // ================= X3: s = prefix(u) ; R = suffix(r)
        float2 u[H], sfull[H], mz[H], r[H], R[H];
        {
            float2 sl[H], Rl[H];
            float utot = 0.f, rtot = 0.f;
            {
                float2 am1[H];
                if (i > 0) {
                    // undo mass preservation on the stored row: the recurrence ran on the raw alpha
                    if constexpr (MASKED) {
                        // the owner of column L-1 swaps the raw value into the staged row for the duration of
                        // its own load (the next iteration reads the same row again as alpha', unpatched)
                        float* ap = reinterpret_cast<float*>(const_cast<unsigned char*>(a_prev_row) +
                                                             (SHIFT ? sh_of(al_lo, i - 1, al_pitch) : 0u)) + last_col;
                        float kept = 0.f;
                        if (k_add >= 0) {
                            kept = *reinterpret_cast<volatile float*>(ap);
                            *reinterpret_cast<volatile float*>(ap) = side_prev_last;
                        }
                        asm volatile("" ::: "memory");
                        load_f(a_prev_row, sh_of(al_lo, i - 1, al_pitch), am1);
                        asm volatile("" ::: "memory");
                        if (k_add >= 0) *reinterpret_cast<volatile float*>(ap) = kept;
                    } else {
                        load_f(a_prev_row, sh_of(al_lo, i - 1, al_pitch), am1);
                        if (mp_last) am1[H - 1].y = side_prev_last;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < H; ++q) am1[q] = f2((j0 + 2 * q == 0) ? 1.0f : 0.0f, 0.0f);
                }
#pragma unroll
                for (int q = 0; q < H; ++q) u[q] = mul2(am1[q], rc[q]);
            }
#pragma unroll
            for (int q = 0; q < H; ++q) {
                utot += u[q].x; sl[q].x = utot;
                utot += u[q].y; sl[q].y = utot;
            }
            if (SOFT) {
                {
                    float2 a_cur[H];
                    load_f(a_cur_row, sh_of(al_lo, i, al_pitch), a_cur);
#pragma unroll
                    for (int q = 0; q < H; ++q) r[q] = mul2(a_cur[q], rD[q]);
                }
                rtot = r[H - 1].y; Rl[H - 1].y = rtot;
                rtot += r[H - 1].x; Rl[H - 1].x = rtot;
#pragma unroll
                for (int q = H - 2; q >= 0; --q) {
                    rtot += r[q].y; Rl[q].y = rtot;
                    rtot += r[q].x; Rl[q].x = rtot;
                }
            }
            // from here on u only feeds gc = -carry*u*1[eps<=cp<=1]: fold the clamp mask in
#pragma unroll
            for (int q = 0; q < H; ++q) u[q] = mul2(u[q], pass[q]);
            float uinc = utot, rinc = rtot;
            if (SOFT) {
                lvl_up_dn<1>(uinc, rinc); lvl_up_dn<2>(uinc, rinc); lvl_up_dn<4>(uinc, rinc);
                lvl_up_dn<8>(uinc, rinc); lvl_up_dn<16>(uinc, rinc);
            } else {
                uinc = wscan_prefix_add(uinc);
            }
            if (lane == 31) xs(B3, 0)[warp] = uinc;
            if (SOFT && lane == 0) xs(B3, 1)[warp] = rinc;
            const float uexc = nb_prev(uinc, 0.f);
            float rexc = 0.f;
            if (SOFT) rexc = nb_next(rinc, 0.f);
            __syncthreads();
            const float2 ubase = f2(xw_off_prefix<NW>(xs(B3, 0), ww, warp, lane, nullptr) + uexc);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                sfull[q] = add2(ubase, sl[q]);
                mz[q] = unit_mask2(mul2(P[q], sfull[q]));
            }
            if (SOFT) {
                const float2 rbase = f2(xw_off_suffix<NW>(xs(B3, 1), ww, warp, lane) + rexc);
#pragma unroll
                for (int q = 0; q < H; ++q) R[q] = add2(rbase, Rl[q]);
            }
        }

        // ================= X4: gr = prefix(gR) ; d/d alpha' = gr/D ; hD = gr*r/D (= -gD)
        float2 gsoft[H], ge1[H], hD[H];
        float g_all = 0.f;
        if (SOFT) {
            float2 grl[H], gB[H];
            float gtot = 0.f;
            if (has_gb) {
                load_f(st + L::kOffGB, sh_of(gb_lo, i, gb_pitch), gB);
            } else {
#pragma unroll
                for (int q = 0; q < H; ++q) gB[q] = f2(0.f);
            }
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float2 gb = mul2(gB[q], unit_mask2(mul2(ex[q], R[q])));
                ge1[q] = mul2(gb, R[q]);
                const float2 gR = mul2(gb, ex[q]);
                gtot += gR.x; grl[q].x = gtot;
                gtot += gR.y; grl[q].y = gtot;
            }
            const float ginc = wscan_prefix_add(gtot);
            if (lane == 31) xs(B4, 0)[warp] = ginc;
            const float gexc = nb_prev(ginc, 0.f);
            __syncthreads();
            const float2 gbase = f2(xw_off_prefix<NW>(xs(B4, 0), ww, warp, lane, mp ? &g_all : nullptr) + gexc);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                gsoft[q] = mul2(add2(gbase, grl[q]), rD[q]);
                hD[q] = mul2(gsoft[q], r[q]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < H; ++q) { gsoft[q] = f2(0.f); ge1[q] = f2(0.f); hD[q] = f2(0.f); }
        }

        // ================= X5: gu = suffix(mz*P*g0) ; suffix(hD)
        // mass preservation: ga_j = g'_j - ok*g'_last (rewritten column: 0); g'_last is formed
        // from block totals every thread already holds
        float2 gA[H];
        float gA_last = 0.f;
        if (has_ga) {
            load_f(st + L::kOffGA, sh_of(ga_lo, i, ga_pitch), gA);
            if (mp) gA_last = reinterpret_cast<const float*>(st + L::kOffGA + (SHIFT ? sh_of(ga_lo, i, ga_pitch) : 0u))[last_col];
        } else {
#pragma unroll
            for (int q = 0; q < H; ++q) gA[q] = f2(0.f);
        }
        if constexpr (DELAYS) if (has_gd) {
            const float2 gd2 = f2(gd_cur), fj = f2((float)j0);
#pragma unroll
            for (int q = 0; q < H; ++q)
                gA[q] = fma2(add2(fj, f2((float)(2 * q + 1), (float)(2 * q + 2))), gd2, gA[q]);
            if (mp) gA_last = __fmaf_rn((float)(last_col + 1), gd_cur, gA_last);
            mask_tail(gA, 0.f);
        }
        float okg = 0.f;
        if (mp) {
            const float ok = (side_sum >= 0.0f && side_sum <= 1.0f) ? 1.0f : 0.0f;
            okg = ok * (gA_last + g_all * rD_last);
        }
        float2 g0[H], gu[H], gEm[H];
        float gEsum = 0.f;
        {
            float2 Al[H], Hl[H];
#pragma unroll
            for (int q = 0; q < H; ++q) g0[q] = add2(gA[q], gsoft[q]);
            if (mp_last) g0[H - 1].y = 0.f;
            const float2 nokg = f2(-okg);
#pragma unroll
            for (int q = 0; q < H; ++q) {
                g0[q] = add2(g0[q], carry[q]);
                if (mp) {
                    const float2 wv = (q == H - 1) ? f2(1.0f, w_lastcol) : f2(1.0f);
                    g0[q] = fma2(wv, nokg, g0[q]);
                }
            }
            if (SHIFT && mp_repl && k_add >= 0) {
                // no-mask rule on a SHIFT row: the output at the REPLACED column does not depend on its raw
                // value, which only feeds the recurrence
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (k == k_add) SIMULST_EL(g0, k) = SIMULST_EL(carry, k);
            }
            float Atot = 0.f, Htot = 0.f;
#pragma unroll
            for (int q = H - 1; q >= 0; --q) {
                const float2 Aq = mul2(mul2(mz[q], P[q]), g0[q]);
                Atot += Aq.y; Al[q].y = Atot;
                Atot += Aq.x; Al[q].x = Atot;
                if (SOFT) {
                    Htot += hD[q].y; Hl[q].y = Htot;
                    Htot += hD[q].x; Hl[q].x = Htot;
                }
            }
            float Ainc = Atot, Hinc = Htot;
            if (SOFT) {
                lvl_dn_dn<1>(Ainc, Hinc); lvl_dn_dn<2>(Ainc, Hinc); lvl_dn_dn<4>(Ainc, Hinc);
                lvl_dn_dn<8>(Ainc, Hinc); lvl_dn_dn<16>(Ainc, Hinc);
            } else {
                Ainc = wscan_suffix_add(Ainc);
            }
            if (lane == 0) {
                xs(B5, 0)[warp] = Ainc;
                if (SOFT) xs(B5, 1)[warp] = Hinc;
            }
            const float Aexc = nb_next(Ainc, 0.f);
            float Hexc = 0.f;
            if (SOFT) Hexc = nb_next(Hinc, 0.f);
            __syncthreads();
            const float2 Abase = f2(xw_off_suffix<NW>(xs(B5, 0), ww, warp, lane) + Aexc);
#pragma unroll
            for (int q = 0; q < H; ++q) gu[q] = add2(Abase, Al[q]);
            if (SOFT) {
                const float2 Hbase = f2(xw_off_suffix<NW>(xs(B5, 1), ww, warp, lane) + Hexc);
                const float2 neg1 = f2(-1.0f);
#pragma unroll
                for (int q = 0; q < H; ++q) {
                    const float2 ge = fma2(add2(Hbase, Hl[q]), neg1, ge1[q]);      // ge1 + suffix(gD)
                    gEm[q] = mul2(ge, exm[q]);
                    gEsum += gEm[q].x;
                    gEsum += gEm[q].y;
                }
            }
        }

        // ================= X6: gL = exclusive suffix of gA ; sum of gEm
        float2 gPk[H], gAl[H];
        float gAtot = 0.f;
        {
            const float2 neg1 = f2(-1.0f);
#pragma unroll
            for (int q = H - 1; q >= 0; --q) {
                gPk[q] = mul2(mul2(mz[q], g0[q]), sfull[q]);
                carry[q] = mul2(gu[q], rc[q]);                       // dL/d alpha_{i-1}
                const float2 gcu = mul2(carry[q], u[q]);             // = -gc * 1[eps<=cp<=1]
                const float2 gcp = fma2(gcu, neg1, mul2(gPk[q], p[q]));
                const float2 gAk = mul2(gcp, cp[q]);
                gAl[q].y = gAtot; gAtot += gAk.y;                    // exclusive local suffix
                gAl[q].x = gAtot; gAtot += gAk.x;
            }
        }
        // row max of the NEXT step's energies (its stage was requested two iterations ago)
        float Emax_next = -INFINITY;
        if (SOFT && i > 0) {
            mbar_wait(&bars[s], parity);
            if (MASKED && fixer) fix_stage(qi + 1);
            Emax_next = row_max(stage0 + s * L::kStage + L::kOffE, i - 1);
        }
        if (MASKED && !SOFT && fixer && i > 0) {
            mbar_wait(&bars[s], parity);
            fix_stage(qi + 1);
        }
        float gAinc = gAtot, ws = gEsum, wmax = Emax_next;
        if (SOFT) {
            lvl_dn_sum_max<1, 16>(gAinc, ws, wmax); lvl_dn_sum_max<2, 8>(gAinc, ws, wmax);
            lvl_dn_sum_max<4, 4>(gAinc, ws, wmax); lvl_dn_sum_max<8, 2>(gAinc, ws, wmax);
            lvl_dn_sum_max<16, 1>(gAinc, ws, wmax);
        } else {
            gAinc = wscan_suffix_add(gAinc);
        }
        if (lane == 0) {
            xs(B6, 0)[warp] = gAinc;
            if (SOFT) { xs(B6, 1)[warp] = ws; xs(B6, 2)[warp] = wmax; }
        }
        const float gAexc = nb_next(gAinc, 0.f);
        // the thread owning the arg-max locates the element now: after the barrier below other
        // warps may already refill this stage
        int k_hit = -1;
        if (SOFT && amax == tid) {
            float2 Er[H];
            load_t(st + L::kOffE, sh_of(e_lo, i, e_pitch), Er);
#pragma unroll
            for (int k = VPT - 1; k >= 0; --k)
                if (SIMULST_EL(Er, k) == m) k_hit = k;
        }
        // 1/((1-p)+eps), recomputed here to keep it out of the registers for the whole step
        float2 rx[H];
#pragma unroll
        for (int q = 0; q < H; ++q) rx[q] = rcp2(add2(fma2(p[q], f2(-1.0f), f2(1.0f)), eps2));
        __syncthreads();
        {
            const float2 gLbase = f2(xw_off_suffix<NW>(xs(B6, 0), ww, warp, lane) + gAexc);
            if (SOFT) {
                m_cur = fxw_max<NW>(xs(B6, 2), lane);
                Emax_cur = Emax_next;
            }
            const float2 neg1 = f2(-1.0f);
            float outp[VPT];
#pragma unroll
            for (int q = 0; q < H; ++q) {
                const float2 gL = add2(gLbase, gAl[q]);
                const float2 o = fma2(mul2(gL, rx[q]), neg1, mul2(gPk[q], cp[q]));
                outp[2 * q] = o.x; outp[2 * q + 1] = o.y;
            }
            if constexpr (MASKED) {
                if (nl > 0) st_row_t<T, VPT, true>(pgp_run, 0, S, true, outp);
                if (nl > 0 && nl < VPT) {
                    // the thread the row ends in: zeros over its columns beyond the row
                    for (int k = nl; k < VPT; ++k) pgp_run[k] = from_f32<T>(0.f);
                }
                pgp_run -= ld_gp;
            } else {
                if (inside) st_row_t<T, VPT, true>(gp_out + (size_t)i * ld_gp, j0, S, true, outp);
            }
        }
        if (SOFT) {
            float oute[VPT];
#pragma unroll
            for (int q = 0; q < H; ++q) { oute[2 * q] = gEm[q].x; oute[2 * q + 1] = gEm[q].y; }
            if (k_hit >= 0) {                   // one thread per row: autograd routes max's gradient to the arg-max
                const float gEall = xw_sum_one<NW>(xs(B6, 1));
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (k == k_hit) oute[k] -= gEall;
            }
            // (columns beyond the row: exp(-inf - m) = 0 makes their energy gradient an exact zero)
            if constexpr (MASKED) {
                if (nl > 0) st_row_t<T, VPT, true>(pge_run, 0, S, true, oute);
                pge_run -= ld_ge;
            } else {
                if (inside) st_row_t<T, VPT, true>(ge_out + (size_t)i * ld_ge, j0, S, true, oute);
            }
        }
        side_sum = side_sum_next;
        side_prev_last = side_prev_next;
        if constexpr (DELAYS) gd_cur = gd_next;
    }
}

template <int THREADS, int VPT, typename T, bool SOFT, bool RAGGED, bool DELAYS, bool MASKED = false, bool SHIFT = false>
int launch_mma_bwd_fast_impl(const MmaParams& prm, cudaStream_t stream) {
    using L = FastLayout<THREADS * VPT, T, SOFT>;
    auto kern = mma_bwd_fast_kernel<THREADS, VPT, T, SOFT, RAGGED, DELAYS, MASKED, SHIFT>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) {
            cudaGetLastError();
            return SIMULST_E_LAUNCH;
        }
        attr_set[dev & 63] = true;
    }
    kern<<<prm.N, THREADS, L::kTotal, stream>>>(prm);
    return check_launch();
}

// returns 1 when the row does not qualify (the caller then uses the generic kernel)
template <int THREADS, int VPT, typename T, bool SOFT>
int launch_mma_bwd_fast(const MmaParams& prm, cudaStream_t stream) {
    constexpr int CAP = THREADS * VPT;
    if constexpr (THREADS / kWarp > kFastMaxWarps || VPT % 4 != 0 || VPT > 12 ||
                  FastLayout<CAP, T, SOFT>::kTotal > 227 * 1024) {
        return 1;
    } else {
        // unmasked rows, TMA staging and 16-byte rows legal, every thread wholly inside or
        // outside the row
        if (prm.shift && prm.S + 16 <= CAP && !(prm.flags & SIMULST_MMA_LEFT_PADDING) &&
            (prm.mask == nullptr || (prm.flags & SIMULST_MMA_RIGHT_PADDING)))
            return launch_mma_bwd_fast_impl<THREADS, VPT, T, SOFT, true, true, true, true>(prm, stream);
        if (!prm.vec_out || !prm.tma || prm.pitched) return 1;
        if (prm.S > CAP || prm.S % VPT != 0) return 1;
        if (prm.mask != nullptr) {
            // masked rows: only when the caller promises a right-padding mask
            if (!(prm.flags & SIMULST_MMA_RIGHT_PADDING) || (prm.flags & SIMULST_MMA_LEFT_PADDING)) return 1;
            return launch_mma_bwd_fast_impl<THREADS, VPT, T, SOFT, true, true, true>(prm, stream);
        }
        if (prm.S != CAP) return launch_mma_bwd_fast_impl<THREADS, VPT, T, SOFT, true, true>(prm, stream);
        return prm.g_delays != nullptr ? launch_mma_bwd_fast_impl<THREADS, VPT, T, SOFT, false, true>(prm, stream)
                                       : launch_mma_bwd_fast_impl<THREADS, VPT, T, SOFT, false, false>(prm, stream);
    }
}

}  // namespace simulst
