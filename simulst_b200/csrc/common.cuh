// Shared device helpers for the simulst_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <type_traits>

#include "../../include/simulst_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "simulst_b200 kernels are written for sm_100a (B200) only"
#endif

namespace simulst {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ dtype conversion
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ------------------------------------------------------------------ vector access
// Load V consecutive elements of type T (as fp32) starting at ptr+idx; `n_valid` of them
// exist (the rest read as `fill`).  `vec_ok` says whether the wide path is legal (all V
// valid and the address is aligned to V*sizeof(T)).
template <typename T, int V>
struct alignas(sizeof(T) * V) Pack { T v[V]; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ base, int n_valid, bool vec_ok,
                                         float fill, float (&out)[V]) {
    if (vec_ok) {
        Pack<T, V> pk = *reinterpret_cast<const Pack<T, V>*>(base);
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = to_f32<T>(pk.v[k]);
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = (k < n_valid) ? to_f32<T>(base[k]) : fill;
    }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ base, int n_valid, bool vec_ok,
                                          const float (&in)[V]) {
    if (vec_ok) {
        Pack<T, V> pk;
#pragma unroll
        for (int k = 0; k < V; ++k) pk.v[k] = from_f32<T>(in[k]);
        *reinterpret_cast<Pack<T, V>*>(base) = pk;
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k)
            if (k < n_valid) base[k] = from_f32<T>(in[k]);
    }
}

// ------------------------------------------------------------------ warp scans
__device__ __forceinline__ float warp_incl_prefix(float v, int lane) {
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        float o = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ float warp_incl_suffix(float v, int lane) {
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        float o = __shfl_down_sync(kFull, v, d);
        if (lane + d < kWarp) v += o;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

// multiplicative variants (exclusive cumprod as a true product scan)
__device__ __forceinline__ float warp_incl_prefix_mul(float v, int lane) {
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        float o = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v *= o;
    }
    return v;
}
// value held by the previous / next lane (identity at the warp edge)
__device__ __forceinline__ float lane_prev(float v, int lane, float ident) {
    float o = __shfl_up_sync(kFull, v, 1);
    return lane == 0 ? ident : o;
}
__device__ __forceinline__ float lane_next(float v, int lane, float ident) {
    float o = __shfl_down_sync(kFull, v, 1);
    return lane == kWarp - 1 ? ident : o;
}

// ------------------------------------------------------------------ fast math (MUFU)
// Relative error <= 2^-22 each; SIMULST_PRECISE_MATH switches to IEEE-rounded versions.
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef SIMULST_PRECISE_MATH
    return __frcp_rn(x);
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
// exp(x) for x <= 0 (soft-attention numerator)
__device__ __forceinline__ float fast_exp(float x) {
#ifdef SIMULST_PRECISE_MATH
    return expf(x);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
#endif
}

// ------------------------------------------------------------------ block exchanges
// A block-wide scan of per-thread totals is: warp scan (shuffles) -> one value per warp in
// shared memory -> ONE __syncthreads -> every thread combines the warp values it needs.
// `slot` arrays are double-buffered by the caller, so no second barrier is required.
//
// Combine step for NW warps: returns {sum of warp totals strictly before `warp`, block total}.
template <int NW>
__device__ __forceinline__ float2 combine_prefix(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return make_float2(0.f, wt[0]);
    } else if constexpr (NW <= 8) {
        float off = 0.f, tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            float v = wt[w];
            if (w < warp) off += v;
            tot += v;
        }
        return make_float2(off, tot);
    } else {
        float v = (lane < NW) ? wt[lane] : 0.f;
        float inc = warp_incl_prefix(v, lane);
        float tot = __shfl_sync(kFull, inc, NW - 1);
        float off = __shfl_sync(kFull, inc - v, warp);
        return make_float2(off, tot);
    }
}
// {sum of warp totals strictly after `warp`, block total}
template <int NW>
__device__ __forceinline__ float2 combine_suffix(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return make_float2(0.f, wt[0]);
    } else if constexpr (NW <= 8) {
        float off = 0.f, tot = 0.f;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            float v = wt[w];
            if (w > warp) off += v;
            tot += v;
        }
        return make_float2(off, tot);
    } else {
        float v = (lane < NW) ? wt[lane] : 0.f;
        float inc = warp_incl_suffix(v, lane);
        float tot = __shfl_sync(kFull, inc, 0);
        float off = __shfl_sync(kFull, inc - v, warp);
        return make_float2(off, tot);
    }
}
// product of warp totals strictly before `warp`
template <int NW>
__device__ __forceinline__ float combine_prefix_mul(const float* __restrict__ wt, int warp, int lane) {
    if constexpr (NW == 1) {
        return 1.0f;
    } else if constexpr (NW <= 8) {
        float off = 1.0f;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) {
            float v = wt[w];
            if (w < warp) off *= v;
        }
        return off;
    } else {
        float v = (lane < NW) ? wt[lane] : 1.0f;
        float inc = warp_incl_prefix_mul(v, lane);
        float exc = lane_prev(inc, lane, 1.0f);
        return __shfl_sync(kFull, exc, warp);
    }
}
template <int NW>
__device__ __forceinline__ float combine_sum(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += wt[w];
        return tot;
    } else {
        return warp_sum((lane < NW) ? wt[lane] : 0.f);
    }
}
template <int NW>
__device__ __forceinline__ float combine_max(const float* __restrict__ wt, int lane) {
    if constexpr (NW <= 8) {
        float m = wt[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, wt[w]);
        return m;
    } else {
        return warp_max((lane < NW) ? wt[lane] : -INFINITY);
    }
}

// ------------------------------------------------------------------ mbarrier / TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// mbar_wait for warps with slack (producers / consumers around a latency-critical warp): back off
// between polls so the polling does not take issue slots from the warp everybody waits for
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done) __nanosleep(40);
    } while (!done);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine;
// SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk async copy shared -> global (bulk-group completion).
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with launch_pdl() may start while its predecessor on the stream is still
// draining: pdl_wait() blocks until every prerequisite grid has completed and its writes are
// visible (no-op for a normal launch); pdl_launch_dependents() lets the successor's CTAs be
// scheduled as soon as SM resources free up.  Every kernel launched this way calls pdl_wait()
// before its first access to global memory.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ status word
__device__ __forceinline__ void flag_status(unsigned* status, unsigned bits) {
    if (status != nullptr && bits != 0u) atomicOr(status, bits);
}
// prob_check (functions.py:9-17): NaN, or outside [0 - 1e-10, 1 + 1e-10] evaluated in fp32
// (1 + 1e-10 rounds to 1.0f).
// One element loaded as raw bits (the consumer converts later, so the load's latency is not
// waited for at the point of the load) and its conversion.  Zero bits are 0.0 in every type.
template <typename T>
__device__ __forceinline__ unsigned ldg_raw(const T* p) {
    unsigned v;
    if constexpr (sizeof(T) == 4) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));      // zero-extended into the 32-bit register
    return v;
}
template <typename T>
__device__ __forceinline__ float raw_to_f32(unsigned v) {
    if constexpr (sizeof(T) == 4) return __uint_as_float(v);
    else if constexpr (std::is_same<T, __nv_bfloat16>::value) return __uint_as_float(v << 16);
    else return __half2float(__ushort_as_half((unsigned short)v));
}
__device__ __forceinline__ unsigned prob_bits(float v) {
    unsigned b = 0u;
    if (v != v) b |= SIMULST_ST_NAN;
    if (v > 1.0f || v < -1e-10f) b |= SIMULST_ST_RANGE;
    return b;
}

// ------------------------------------------------------------------ host side
struct LaunchCounter {
    static std::atomic<long long>& value() {
        static std::atomic<long long> v{0};
        return v;
    }
};

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in check_launch()
}

inline int check_launch() {
    LaunchCounter::value().fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SIMULST_OK : SIMULST_E_LAUNCH;
}

inline bool valid_dtype(int d) { return d == SIMULST_F32 || d == SIMULST_BF16 || d == SIMULST_F16; }
inline size_t dtype_size(int d) { return d == SIMULST_F32 ? 4 : 2; }

}  // namespace simulst
