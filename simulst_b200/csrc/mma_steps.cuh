// Per-thread pieces of one MMA target step, shared VERBATIM by the pipelined forward and
// backward kernels: the backward recomputes the forward's scans and must reproduce them bit
// for bit (its clamp masks compare the recomputed values with 0 / 1 / eps).
//
// Math: SURVEY Appendix A.1 / A.3 (reference codebase/utils/monotonic_attention.py:40-69 and
// :103-150, codebase/utils/functions.py:20-66).  A thread owns VPT consecutive source
// positions held as VPT/2 float2 pairs.
#pragma once

#include "mma_scan.cuh"

namespace simulst {

// x_k = (1 - p_k) + eps ;  cpre_k = prod_{q<k} x_q (thread-local exclusive product) ; returns
// the thread total.  (1-p) and +eps stay two roundings: p = 1 must give exactly eps.
template <int VPT>
__device__ __forceinline__ float local_cumprod(const float2 (&p)[VPT / 2], float eps, float2 (&cpre)[VPT / 2]) {
    float2 x[VPT / 2];
    const float2 one = f2(1.0f), neg = f2(-1.0f), e2 = f2(eps);
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) x[q] = add2(fma2(p[q], neg, one), e2);
    float xt = 1.0f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        SIMULST_EL(cpre, k) = xt;
        xt *= SIMULST_EL(x, k);
    }
    return xt;
}

// cp = cbase * cpre ;  rc = 1 / clamp(cp, eps, 1) ;  P = p * cp
template <int VPT>
__device__ __forceinline__ void finish_cumprod(float cbase, const float2 (&cpre)[VPT / 2], const float2 (&p)[VPT / 2],
                                               float eps, float2 (&cp)[VPT / 2], float2 (&rc)[VPT / 2],
                                               float2 (&P)[VPT / 2]) {
    const float2 cb = f2(cbase);
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) {
        cp[q] = mul2(cb, cpre[q]);
        rc[q] = make_float2(fast_rcp(fminf(fmaxf(cp[q].x, eps), 1.0f)), fast_rcp(fminf(fmaxf(cp[q].y, eps), 1.0f)));
        P[q] = mul2(p[q], cp[q]);
    }
}

// exm_k = exp(E_k - m) ;  ex_k = exm_k + eps ;  Dl_k = sum_{q<=k} ex_q (thread-local inclusive
// prefix) ; returns the thread total.  exp is ex2((E - m) * log2e) evaluated as one FMA.
template <int VPT, bool KEEP_EXM>
__device__ __forceinline__ float local_exp_prefix(const float2 (&E)[VPT / 2], float m, float eps,
                                                  float2 (&exm)[VPT / 2], float2 (&ex)[VPT / 2], float2 (&Dl)[VPT / 2]) {
    const float2 l2 = f2(kLog2e), mm = f2(-m * kLog2e), e2 = f2(eps);
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) {
        const float2 a = fma2(E[q], l2, mm);
        const float2 t = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        if (KEEP_EXM) exm[q] = t;
        ex[q] = add2(t, e2);
    }
    float et = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        et += SIMULST_EL(ex, k);
        SIMULST_EL(Dl, k) = et;
    }
    return et;
}

// u_k = a_k * rc_k ;  sl_k = sum_{q<=k} u_q (FMA chain) ; returns the thread total
template <int VPT>
__device__ __forceinline__ float local_u_prefix(const float2 (&a)[VPT / 2], const float2 (&rc)[VPT / 2],
                                                float2 (&sl)[VPT / 2]) {
    float ut = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        ut = fmaf(SIMULST_EL(a, k), SIMULST_EL(rc, k), ut);
        SIMULST_EL(sl, k) = ut;
    }
    return ut;
}

// z_k = P_k * (ubase + sl_k)   (alpha before the clamp)
template <int VPT>
__device__ __forceinline__ void finish_u_prefix(float ubase, const float2 (&sl)[VPT / 2], const float2 (&P)[VPT / 2],
                                                float2 (&sfull)[VPT / 2], float2 (&z)[VPT / 2]) {
    const float2 ub = f2(ubase);
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) {
        sfull[q] = add2(ub, sl[q]);
        z[q] = mul2(P[q], sfull[q]);
    }
}

// 1 / D_k with D_k = (eps + ebase) + Dl_k
template <int VPT>
__device__ __forceinline__ void finish_exp_prefix(float ebase, float eps, const float2 (&Dl)[VPT / 2],
                                                  float2 (&rD)[VPT / 2]) {
    const float2 db = f2(eps + ebase);
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) rD[q] = rcp2(add2(db, Dl[q]));
}

// r_k = a_k * rD_k ;  Rl_k = sum_{q>=k} r_q (thread-local inclusive suffix, FMA chain).
// Returns Rl_0 (= thread total).
template <int VPT>
__device__ __forceinline__ float local_r_suffix(const float2 (&a)[VPT / 2], const float2 (&rD)[VPT / 2],
                                                float2 (&Rl)[VPT / 2]) {
    float rt = 0.f;
#pragma unroll
    for (int k = VPT - 1; k >= 0; --k) {
        rt = fmaf(SIMULST_EL(a, k), SIMULST_EL(rD, k), rt);
        SIMULST_EL(Rl, k) = rt;
    }
    return rt;
}

}  // namespace simulst
