// Warp-scan and packed-math primitives of the pipelined MMA kernels (sm_100a).
//
// * Scans use `shfl.sync` with its predicate output (source lane in range), so one scan step
//   is SHFL + one predicated FADD/FMUL -- no lane compare, no select.
// * f32x2 helpers map to the Blackwell packed FP32 instructions (FADD2 / FMUL2 / FFMA2): the
//   element-wise part of a step issues one instruction per two source positions.
// * Everything here is shared by the forward and the backward kernel so that the scans the
//   backward recomputes are bit-identical to the forward's (the clamp masks depend on it).
#pragma once

#include <type_traits>

#include "common.cuh"

namespace simulst {

// ------------------------------------------------------------------ packed fp32
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 min2(float2 a, float b) { return make_float2(fminf(a.x, b), fminf(a.y, b)); }
__device__ __forceinline__ float2 rcp2(float2 a) { return make_float2(fast_rcp(a.x), fast_rcp(a.y)); }

// element k of an array of float2 pairs (k is a compile-time constant after unrolling)
#define SIMULST_EL(arr, k) (((k)&1) ? (arr)[(k) >> 1].y : (arr)[(k) >> 1].x)

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef SIMULST_PRECISE_MATH
    return exp2f(x);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
constexpr float kLog2e = 1.4426950408889634f;

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0u;
}

// ------------------------------------------------------------------ warp scans (predicated shuffles)
// The asm statements are `volatile`: a pure asm may be sunk by the compiler into code that only
// some lanes execute (it then wraps every shuffle in WARPSYNC.COLLECTIVE / ENDCOLLECTIVE).
template <int D>
__device__ __forceinline__ void shfl_up_add(float& v) {
    asm volatile("{\n\t.reg .f32 t;\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 t|q, %0, %1, 0, 0xffffffff;\n\t"
        "@q add.rn.f32 %0, %0, t;\n\t}"
        : "+f"(v)
        : "n"(D));
}
template <int D>
__device__ __forceinline__ void shfl_up_mul(float& v) {
    asm volatile("{\n\t.reg .f32 t;\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 t|q, %0, %1, 0, 0xffffffff;\n\t"
        "@q mul.rn.f32 %0, %0, t;\n\t}"
        : "+f"(v)
        : "n"(D));
}
template <int D>
__device__ __forceinline__ void shfl_down_add(float& v) {
    asm volatile("{\n\t.reg .f32 t;\n\t.reg .pred q;\n\t"
        "shfl.sync.down.b32 t|q, %0, %1, 31, 0xffffffff;\n\t"
        "@q add.rn.f32 %0, %0, t;\n\t}"
        : "+f"(v)
        : "n"(D));
}
// inclusive scans over the 32 lanes
__device__ __forceinline__ float wscan_prefix_add(float v) {
    shfl_up_add<1>(v); shfl_up_add<2>(v); shfl_up_add<4>(v); shfl_up_add<8>(v); shfl_up_add<16>(v);
    return v;
}
__device__ __forceinline__ float wscan_prefix_mul(float v) {
    shfl_up_mul<1>(v); shfl_up_mul<2>(v); shfl_up_mul<4>(v); shfl_up_mul<8>(v); shfl_up_mul<16>(v);
    return v;
}
__device__ __forceinline__ float wscan_suffix_add(float v) {
    shfl_down_add<1>(v); shfl_down_add<2>(v); shfl_down_add<4>(v); shfl_down_add<8>(v); shfl_down_add<16>(v);
    return v;
}
// Fused scans of the pipelined kernels: several independent inclusive scans advance level by
// level in one asm statement, so their shuffle latencies overlap by construction.
//   x: prefix product   e, u: prefix sums   r: suffix sum
template <int D>
__device__ __forceinline__ void scan_level_xeur(float& x, float& e, float& u, float& r) {
    asm volatile("{\n\t.reg .f32 t0, t1, t2, t3;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t1, %1, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t2, %2, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t3|q1, %3, %4, 31, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "@q0 add.rn.f32 %1, %1, t1;\n\t"
        "@q0 add.rn.f32 %2, %2, t2;\n\t"
        "@q1 add.rn.f32 %3, %3, t3;\n\t}"
        : "+f"(x), "+f"(e), "+f"(u), "+f"(r)
        : "n"(D));
}
__device__ __forceinline__ void wscan_xeur(float& x, float& e, float& u, float& r) {
    scan_level_xeur<1>(x, e, u, r); scan_level_xeur<2>(x, e, u, r); scan_level_xeur<4>(x, e, u, r);
    scan_level_xeur<8>(x, e, u, r); scan_level_xeur<16>(x, e, u, r);
}
template <int D>
__device__ __forceinline__ void scan_level_xu(float& x, float& u) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t1, %1, %2, 0, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "@q0 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(x), "+f"(u)
        : "n"(D));
}
__device__ __forceinline__ void wscan_xu(float& x, float& u) {
    scan_level_xu<1>(x, u); scan_level_xu<2>(x, u); scan_level_xu<4>(x, u); scan_level_xu<8>(x, u); scan_level_xu<16>(x, u);
}
// neighbours of the four inclusive results: previous lane for x / e / u (identity 1 / 0 / 0 at
// lane 0), next lane for r (0 at lane 31)
__device__ __forceinline__ void wneigh_xeur(float x, float e, float u, float r, float& xp, float& ep, float& up, float& rn) {
    asm volatile("{\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 %0|q0, %4, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 %1, %5, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 %2, %6, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %3|q1, %7, 1, 31, 0xffffffff;\n\t"
        "@!q0 mov.f32 %0, 0f3F800000;\n\t"
        "@!q0 mov.f32 %1, 0f00000000;\n\t"
        "@!q0 mov.f32 %2, 0f00000000;\n\t"
        "@!q1 mov.f32 %3, 0f00000000;\n\t}"
        : "=&f"(xp), "=&f"(ep), "=&f"(up), "=&f"(rn)
        : "f"(x), "f"(e), "f"(u), "f"(r));
}
// ---- backward kernel, barrier 1: x prefix product, e prefix sum, q prefix sum (only its
// total at lane 31 is used), g suffix sum
template <int D>
__device__ __forceinline__ void scan_level_b1(float& x, float& e, float& q, float& g) {
    asm volatile("{\n\t.reg .f32 t0, t1, t2, t3;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t1, %1, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 t2, %2, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t3|q1, %3, %4, 31, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "@q0 add.rn.f32 %1, %1, t1;\n\t"
        "@q0 add.rn.f32 %2, %2, t2;\n\t"
        "@q1 add.rn.f32 %3, %3, t3;\n\t}"
        : "+f"(x), "+f"(e), "+f"(q), "+f"(g)
        : "n"(D));
}
__device__ __forceinline__ void wscan_b1(float& x, float& e, float& q, float& g) {
    scan_level_b1<1>(x, e, q, g); scan_level_b1<2>(x, e, q, g); scan_level_b1<4>(x, e, q, g);
    scan_level_b1<8>(x, e, q, g); scan_level_b1<16>(x, e, q, g);
}
// neighbours: previous lane of x (1 at lane 0) and e (0), next lane of g (0 at lane 31)
__device__ __forceinline__ void wneigh_b1(float x, float e, float g, float& xp, float& ep, float& gn) {
    asm volatile("{\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 %0|q0, %3, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 %1, %4, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %2|q1, %5, 1, 31, 0xffffffff;\n\t"
        "@!q0 mov.f32 %0, 0f3F800000;\n\t"
        "@!q0 mov.f32 %1, 0f00000000;\n\t"
        "@!q1 mov.f32 %2, 0f00000000;\n\t}"
        : "=&f"(xp), "=&f"(ep), "=&f"(gn)
        : "f"(x), "f"(e), "f"(g));
}
// ---- backward kernel, barrier 2: u prefix sum; R, W, L suffix sums
template <int D>
__device__ __forceinline__ void scan_level_b2(float& u, float& r, float& w, float& l) {
    asm volatile("{\n\t.reg .f32 t0, t1, t2, t3;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %4, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t1|q1, %1, %4, 31, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t2, %2, %4, 31, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t3, %3, %4, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "@q1 add.rn.f32 %1, %1, t1;\n\t"
        "@q1 add.rn.f32 %2, %2, t2;\n\t"
        "@q1 add.rn.f32 %3, %3, t3;\n\t}"
        : "+f"(u), "+f"(r), "+f"(w), "+f"(l)
        : "n"(D));
}
__device__ __forceinline__ void wscan_b2(float& u, float& r, float& w, float& l) {
    scan_level_b2<1>(u, r, w, l); scan_level_b2<2>(u, r, w, l); scan_level_b2<4>(u, r, w, l);
    scan_level_b2<8>(u, r, w, l); scan_level_b2<16>(u, r, w, l);
}
__device__ __forceinline__ void wneigh_b2(float u, float r, float w, float l, float& up, float& rn, float& wn, float& ln) {
    asm volatile("{\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 %0|q0, %4, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %1|q1, %5, 1, 31, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %2, %6, 1, 31, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %3, %7, 1, 31, 0xffffffff;\n\t"
        "@!q0 mov.f32 %0, 0f00000000;\n\t"
        "@!q1 mov.f32 %1, 0f00000000;\n\t"
        "@!q1 mov.f32 %2, 0f00000000;\n\t"
        "@!q1 mov.f32 %3, 0f00000000;\n\t}"
        : "=&f"(up), "=&f"(rn), "=&f"(wn), "=&f"(ln)
        : "f"(u), "f"(r), "f"(w), "f"(l));
}
// ---- one prefix sum + one suffix sum (backward barrier 3: gr / V; hard-mode barriers)
template <int D>
__device__ __forceinline__ void scan_level_ud(float& u, float& d) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t1|q1, %1, %2, 31, 0xffffffff;\n\t"
        "@q0 add.rn.f32 %0, %0, t0;\n\t"
        "@q1 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(u), "+f"(d)
        : "n"(D));
}
__device__ __forceinline__ void wscan_ud(float& u, float& d) {
    scan_level_ud<1>(u, d); scan_level_ud<2>(u, d); scan_level_ud<4>(u, d); scan_level_ud<8>(u, d); scan_level_ud<16>(u, d);
}
// prefix product + suffix sum (hard-mode backward barrier 1)
template <int D>
__device__ __forceinline__ void scan_level_xd(float& x, float& d) {
    asm volatile("{\n\t.reg .f32 t0, t1;\n\t.reg .pred q0, q1;\n\t"
        "shfl.sync.up.b32 t0|q0, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.down.b32 t1|q1, %1, %2, 31, 0xffffffff;\n\t"
        "@q0 mul.rn.f32 %0, %0, t0;\n\t"
        "@q1 add.rn.f32 %1, %1, t1;\n\t}"
        : "+f"(x), "+f"(d)
        : "n"(D));
}
__device__ __forceinline__ void wscan_xd(float& x, float& d) {
    scan_level_xd<1>(x, d); scan_level_xd<2>(x, d); scan_level_xd<4>(x, d); scan_level_xd<8>(x, d); scan_level_xd<16>(x, d);
}
// value of the previous / next lane, `ident` at the warp edge
__device__ __forceinline__ float wprev(float v, float ident) {
    float o;
    asm volatile("{\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 %0|q, %1, 1, 0, 0xffffffff;\n\t"
        "@!q mov.f32 %0, %2;\n\t}"
        : "=&f"(o)
        : "f"(v), "f"(ident));
    return o;
}
__device__ __forceinline__ float wnext(float v, float ident) {
    float o;
    asm volatile("{\n\t.reg .pred q;\n\t"
        "shfl.sync.down.b32 %0|q, %1, 1, 31, 0xffffffff;\n\t"
        "@!q mov.f32 %0, %2;\n\t}"
        : "=&f"(o)
        : "f"(v), "f"(ident));
    return o;
}
// max over the warp through the integer REDUX unit (order-preserving float <-> int map)
__device__ __forceinline__ float wmax_redux(float v) {
    int k = __float_as_int(v);
    k ^= (k >> 31) & 0x7fffffff;
    k = __reduce_max_sync(kFull, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
}

// ------------------------------------------------------------------ staged-row reads
// Read VPT consecutive elements of a staged row from shared memory as float2 pairs.
// `umax` accumulates the unsigned maximum of the raw bit patterns (16-bit types: per half of
// a packed word): a value > bits(1.0) means "negative, > 1 or NaN" -- the cheap first-level
// prob_check of a FULL row; the exact classification runs only when it trips.
template <typename T> struct RawOne;
template <> struct RawOne<float> { static constexpr unsigned bits = 0x3F800000u; };
template <> struct RawOne<__nv_bfloat16> { static constexpr unsigned bits = 0x3F80u; };
template <> struct RawOne<__half> { static constexpr unsigned bits = 0x3C00u; };

template <typename T, int VPT, bool CHECK>
__device__ __forceinline__ void lds_row2(const T* __restrict__ src, float2 (&out)[VPT / 2], unsigned& umax) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < VPT / 4; ++q) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + 4 * q);
            if (CHECK) umax = max(max(umax, w.x), max(max(w.y, w.z), w.w));
            out[2 * q] = make_float2(__uint_as_float(w.x), __uint_as_float(w.y));
            out[2 * q + 1] = make_float2(__uint_as_float(w.z), __uint_as_float(w.w));
        }
    } else {
        constexpr int WORDS = VPT / 2;
        unsigned w[WORDS];
        if constexpr (WORDS % 4 == 0) {
#pragma unroll
            for (int q = 0; q < WORDS / 4; ++q) {
                const uint4 t = *reinterpret_cast<const uint4*>(src + 8 * q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < WORDS / 2; ++q) {
                const uint2 t = *reinterpret_cast<const uint2*>(src + 4 * q);
                w[2 * q] = t.x; w[2 * q + 1] = t.y;
            }
        }
#pragma unroll
        for (int q = 0; q < WORDS; ++q) {
            if (CHECK) umax = __vmaxu2(umax, w[q]);
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                out[q] = make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u));
            } else {
                out[q] = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
            }
        }
    }
}
template <typename T>
__device__ __forceinline__ bool umax_trips(unsigned umax) {
    if constexpr (sizeof(T) == 4) return umax > RawOne<T>::bits;
    else return (umax & 0xffffu) > RawOne<T>::bits || (umax >> 16) > RawOne<T>::bits;
}

// fp32 row stores from float2 pairs
template <int VPT, bool FULL>
__device__ __forceinline__ void st_row2_f32(float* __restrict__ row, int j0, int S, bool vec,
                                            const float2 (&v)[VPT / 2]) {
#pragma unroll
    for (int q = 0; q < VPT / 4; ++q) {
        const int j = j0 + 4 * q;
        const float4 t = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        if constexpr (FULL) {
            *reinterpret_cast<float4*>(row + j) = t;
        } else if (vec) {
            if (j < S) *reinterpret_cast<float4*>(row + j) = t;
        } else {
            if (j < S) row[j] = t.x;
            if (j + 1 < S) row[j + 1] = t.y;
            if (j + 2 < S) row[j + 2] = t.z;
            if (j + 3 < S) row[j + 3] = t.w;
        }
    }
}

// ------------------------------------------------------------------ cross-warp combines
// Each warp publishes one value per scan in shared memory; after the barrier every thread
// folds the values of the warps before (prefix) / after (suffix) its own.
template <int NW>
__device__ __forceinline__ float xw_prefix_add(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return 0.f;
    } else if constexpr (NW <= 8) {
        float off = 0.f;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) {
            const float v = wt[w];
            if (w < warp) off += v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane] : 0.f;
        const float inc = wscan_prefix_add(v);
        return __shfl_sync(kFull, wprev(inc, 0.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float xw_prefix_mul(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return 1.f;
    } else if constexpr (NW <= 8) {
        float off = 1.f;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) {
            const float v = wt[w];
            if (w < warp) off *= v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane] : 1.f;
        const float inc = wscan_prefix_mul(v);
        return __shfl_sync(kFull, wprev(inc, 1.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float xw_suffix_add(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return 0.f;
    } else if constexpr (NW <= 8) {
        float off = 0.f;
#pragma unroll
        for (int w = NW - 1; w > 0; --w) {
            const float v = wt[w];
            if (w > warp) off += v;
        }
        return off;
    } else {
        const float v = (lane < NW) ? wt[lane] : 0.f;
        const float inc = wscan_suffix_add(v);
        return __shfl_sync(kFull, wnext(inc, 0.f), warp);
    }
}
template <int NW>
__device__ __forceinline__ float xw_sum(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float tot = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) tot += wt[w];
        return tot;
    } else {
        return warp_sum((lane < NW) ? wt[lane] : 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float xw_max(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float m = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, wt[w]);
        return m;
    } else {
        return wmax_redux((lane < NW) ? wt[lane] : -INFINITY);
    }
}

}  // namespace simulst
