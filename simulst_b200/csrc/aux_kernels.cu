// Stand-alone row kernels: the reference functions called one by one (not through the fused
// training kernel).  One CTA per (n, t) row, row staged in shared memory as fp32, block scans
// over shared memory.  These mirror the reference formulas literally (IEEE division, libm-grade
// expf/logf) -- they are utility entry points, not the bandwidth-critical path.
//
//   simulst_soft_attention_fwd/bwd    codebase/utils/monotonic_attention.py:79-152
//   simulst_mass_preservation_fwd/bwd codebase/utils/monotonic_attention.py:155-197
//   simulst_moving_sum                codebase/utils/functions.py:69-125
//   simulst_exclusive_cumprod         codebase/utils/functions.py:20-66
//   simulst_p_choose                  codebase/utils/p_choose_strategy.py:56-76
#include "common.cuh"

namespace simulst {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / kWarp;

// In-place inclusive scan of a[0..S) in shared memory; scratch has kRowThreads + 2 floats.
// Returns the total.  All threads of the CTA must call it.
template <bool SUFFIX>
__device__ float block_scan_smem(float* a, int S, float* scratch) {
    const int tid = threadIdx.x;
    const int per = (S + kRowThreads - 1) / kRowThreads;
    const int lo = min(tid * per, S), hi = min(lo + per, S);
    float run = 0.f;
    if (!SUFFIX) {
        for (int j = lo; j < hi; ++j) { run += a[j]; a[j] = run; }
    } else {
        for (int j = hi - 1; j >= lo; --j) { run += a[j]; a[j] = run; }
    }
    scratch[tid] = run;
    __syncthreads();
    if (tid < kWarp) {
        // 256 partials: each lane owns 8 consecutive ones
        float part[kRowThreads / kWarp];
        float tot = 0.f;
        if (!SUFFIX) {
#pragma unroll
            for (int q = 0; q < kRowThreads / kWarp; ++q) { part[q] = tot; tot += scratch[tid * 8 + q]; }
            const float inc = warp_incl_prefix(tot, tid);
            const float exc = lane_prev(inc, tid, 0.f);
#pragma unroll
            for (int q = 0; q < kRowThreads / kWarp; ++q) scratch[tid * 8 + q] = exc + part[q];
            if (tid == kWarp - 1) scratch[kRowThreads] = inc;
        } else {
#pragma unroll
            for (int q = kRowThreads / kWarp - 1; q >= 0; --q) { part[q] = tot; tot += scratch[tid * 8 + q]; }
            const float inc = warp_incl_suffix(tot, tid);
            const float exc = lane_next(inc, tid, 0.f);
#pragma unroll
            for (int q = 0; q < kRowThreads / kWarp; ++q) scratch[tid * 8 + q] = exc + part[q];
            if (tid == 0) scratch[kRowThreads] = inc;
        }
    }
    __syncthreads();
    const float off = scratch[tid];
    const float total = scratch[kRowThreads];
    for (int j = lo; j < hi; ++j) a[j] += off;
    __syncthreads();
    return total;
}

__device__ float block_reduce_max(float v, float* scratch) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = scratch[0];
#pragma unroll
    for (int w = 1; w < kRowWarps; ++w) m = fmaxf(m, scratch[w]);
    __syncthreads();
    return m;
}
__device__ float block_reduce_sum(float v, float* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) m += scratch[w];
    __syncthreads();
    return m;
}
__device__ int block_reduce_min_int(int v, int* scratch) {
    v = __reduce_min_sync(kFull, v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int m = scratch[0];
#pragma unroll
    for (int w = 1; w < kRowWarps; ++w) m = min(m, scratch[w]);
    __syncthreads();
    return m;
}

// out[j] = sum of in[j-back .. j+fwd] clipped to the row (direct summation like conv1d).
__device__ void window_sum(const float* in, float* out, int S, int back, int fwd) {
    for (int j = threadIdx.x; j < S; j += kRowThreads) {
        float acc = 0.f;
        const int lo = max(0, j - back), hi = min(S - 1, j + fwd);
        for (int q = lo; q <= hi; ++q) acc += in[q];
        out[j] = acc;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------- soft attention
// smem: 5 rows of S floats + scratch
template <typename TA, typename TE, bool BWD>
__global__ void __launch_bounds__(kRowThreads)
soft_attention_kernel(const TA* __restrict__ alpha, const TE* __restrict__ energy,
                      const uint8_t* __restrict__ mask, TA* __restrict__ beta,
                      const TA* __restrict__ g_beta, TA* __restrict__ g_alpha, TE* __restrict__ g_energy,
                      int T, int S, float eps, int chunk, float fill, unsigned* status) {
    extern __shared__ float sm[];
    float* e = sm;               // exp(E - m)   (eps added on use)
    float* D = sm + S;           // eps + cumsum / window of (e + eps)
    float* r = sm + 2 * S;       // alpha / D
    float* w0 = sm + 3 * S;      // work
    float* w1 = sm + 4 * S;      // work
    float* tmp = sm + 5 * S;     // window scratch (chunkwise only)
    float* scratch = sm + 6 * S;
    (void)tmp;
    const size_t row = blockIdx.x;
    const int n = (int)(row / T);
    const TA* a_row = alpha + row * S;
    const TE* e_row = energy + row * S;
    const uint8_t* m_row = mask ? mask + (size_t)n * S : nullptr;
    const int tid = threadIdx.x;

    float mx = -INFINITY;
    unsigned bits = 0u;
    int first_max = 0x7fffffff;
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        const float a = pad ? 0.f : to_f32<TA>(a_row[j]);
        const float en = pad ? fill : to_f32<TE>(e_row[j]);
        bits |= prob_bits(a);
        w0[j] = a;          // masked alpha
        w1[j] = en;         // masked energy
        mx = fmaxf(mx, en);
    }
    mx = block_reduce_max(mx, scratch);
    for (int j = tid; j < S; j += kRowThreads) {
        if (BWD && w1[j] == mx) first_max = min(first_max, j);
        const float ex = expf(w1[j] - mx);
        e[j] = ex;
        D[j] = ex + eps;
    }
    __syncthreads();
    if (chunk > 0) {
        for (int j = tid; j < S; j += kRowThreads) tmp[j] = D[j];
        __syncthreads();
        window_sum(tmp, D, S, chunk - 1, 0);
        for (int j = tid; j < S; j += kRowThreads) D[j] += eps;
    } else {
        block_scan_smem<false>(D, S, scratch);
        for (int j = tid; j < S; j += kRowThreads) D[j] = eps + D[j];
    }
    __syncthreads();
    for (int j = tid; j < S; j += kRowThreads) r[j] = w0[j] / D[j];
    __syncthreads();
    // R into w1 (energy no longer needed)
    if (chunk > 0) {
        window_sum(r, w1, S, 0, chunk - 1);
    } else {
        for (int j = tid; j < S; j += kRowThreads) w1[j] = r[j];
        __syncthreads();
        block_scan_smem<true>(w1, S, scratch);
    }
    if (!BWD) {
        TA* b_row = beta + row * S;
        for (int j = tid; j < S; j += kRowThreads) {
            const bool pad = m_row && m_row[j];
            float b = pad ? 0.f : (e[j] + eps) * w1[j];
            b = to_f32<TA>(from_f32<TA>(b));          // cast to alpha's dtype before the clamp (:146-148)
            if (b != b) bits |= SIMULST_ST_NAN;
            b_row[j] = from_f32<TA>(fminf(fmaxf(b, 0.f), 1.f));
        }
        flag_status(status, bits);
        return;
    }
    // ---------------- backward
    const TA* gb_row = g_beta + row * S;
    // gb -> ge1 = gb*R (kept in w0 after alpha is consumed? alpha still needed: r holds alpha/D)
    // stage: w0 <- gR = gb*e ; keep ge1 in registers-free form by recomputing gb later
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        float b = pad ? 0.f : (e[j] + eps) * w1[j];
        b = to_f32<TA>(from_f32<TA>(b));
        const float gb = (!pad && b >= 0.f && b <= 1.f) ? to_f32<TA>(gb_row[j]) : 0.f;
        w0[j] = gb * (e[j] + eps);   // gR
        w1[j] = gb * w1[j];          // ge1 = gb * R
    }
    __syncthreads();
    // gr = prefix / look-back window of gR  (in place in w0 for prefix; via D-sized temp for window)
    float* gr = w0;
    if (chunk > 0) {
        window_sum(w0, tmp, S, chunk - 1, 0);
        for (int j = tid; j < S; j += kRowThreads) w0[j] = tmp[j];
        __syncthreads();
    } else {
        block_scan_smem<false>(w0, S, scratch);
    }
    // g_alpha = gr / D ; gD = -gr * r / D
    TA* ga_row = g_alpha + row * S;
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        const float gq = gr[j] / D[j];
        ga_row[j] = from_f32<TA>(pad ? 0.f : gq);
        w0[j] = -gq * r[j];          // gD
    }
    __syncthreads();
    if (chunk > 0) {
        window_sum(w0, tmp, S, 0, chunk - 1);
        for (int j = tid; j < S; j += kRowThreads) w0[j] = tmp[j];
        __syncthreads();
    } else {
        block_scan_smem<true>(w0, S, scratch);
    }
    float gsum = 0.f;
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        const float ge = w1[j] + w0[j];
        const float gEm = ge * e[j];                // d(exp(E-m)+eps)/dE = exp(E-m)
        w1[j] = pad ? 0.f : gEm;
        gsum += gEm;                                 // max() sees the masked-filled tensor: all columns
    }
    gsum = block_reduce_sum(gsum, scratch);
    const int amax = block_reduce_min_int(first_max, reinterpret_cast<int*>(scratch));
    TE* ge_row = g_energy + row * S;
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        float v = w1[j];
        if (j == amax) v -= gsum;
        ge_row[j] = from_f32<TE>(pad ? 0.f : v);
    }
}

// ---------------------------------------------------------------------------- mass preservation
__global__ void __launch_bounds__(kRowThreads)
mass_preservation_fwd_kernel(float* __restrict__ alpha, const uint8_t* __restrict__ mask,
                             float* __restrict__ side, int T, int S, unsigned flags, unsigned* status) {
    __shared__ float scratch[kRowWarps + 2];
    __shared__ int iscratch[kRowWarps + 2];
    const size_t row = blockIdx.x;
    const int n = (int)(row / T);
    float* a = alpha + row * S;
    const uint8_t* m_row = mask ? mask + (size_t)n * S : nullptr;
    const bool add_mode = mask != nullptr && !(flags & SIMULST_MMA_LEFT_PADDING);
    const int tid = threadIdx.x;
    int live = 0;
    unsigned bits = 0u;
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        bits |= prob_bits(a[j]);
        if (pad) a[j] = 0.f; else ++live;
    }
    int last = S - 1;
    if (add_mode) {
        live = __reduce_add_sync(kFull, live);
        if ((tid & 31) == 0) iscratch[tid >> 5] = live;
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < kRowWarps; ++w) tot += iscratch[w];
        last = tot - 1;
    }
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < S; j += kRowThreads)
        if (add_mode || j != last) sum += a[j];
    sum = block_reduce_sum(sum, scratch);
    if (tid == 0 && last >= 0) {
        const float resid = 1.0f - fminf(fmaxf(sum, 0.f), 1.f);
        const float old = a[last];
        a[last] = add_mode ? old + resid : resid;
        if (side) { side[row * 2] = old; side[row * 2 + 1] = sum; }
    }
    flag_status(status, bits);
}

__global__ void __launch_bounds__(kRowThreads)
mass_preservation_bwd_kernel(const float* __restrict__ g_in, const uint8_t* __restrict__ mask,
                             const float* __restrict__ side, float* __restrict__ g_out,
                             int T, int S, unsigned flags) {
    __shared__ int iscratch[kRowWarps + 2];
    const size_t row = blockIdx.x;
    const int n = (int)(row / T);
    const uint8_t* m_row = mask ? mask + (size_t)n * S : nullptr;
    const bool add_mode = mask != nullptr && !(flags & SIMULST_MMA_LEFT_PADDING);
    const int tid = threadIdx.x;
    int last = S - 1;
    if (add_mode) {
        int live = 0;
        for (int j = tid; j < S; j += kRowThreads) live += (m_row[j] == 0);
        live = __reduce_add_sync(kFull, live);
        if ((tid & 31) == 0) iscratch[tid >> 5] = live;
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < kRowWarps; ++w) tot += iscratch[w];
        last = tot - 1;
    }
    const float sum = side[row * 2 + 1];
    const float ok = (sum >= 0.f && sum <= 1.f) ? 1.f : 0.f;
    const float glast = last >= 0 ? g_in[row * S + last] : 0.f;
    __syncthreads();        // g_out may alias g_in: everyone has read glast
    for (int j = tid; j < S; j += kRowThreads) {
        const bool pad = m_row && m_row[j];
        float g = g_in[row * S + j] - ok * glast;
        if (!add_mode && j == last) g = 0.f;
        g_out[row * S + j] = pad ? 0.f : g;
    }
}

// ---------------------------------------------------------------------------- moving_sum
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
moving_sum_kernel(const T* __restrict__ x, T* __restrict__ out, int S, int start_idx, int end_idx) {
    extern __shared__ float sm[];
    const size_t row = blockIdx.x;
    for (int j = threadIdx.x; j < S; j += kRowThreads) sm[j] = to_f32<T>(x[row * S + j]);
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += kRowThreads) {
        float acc = 0.f;
        const int lo = max(0, j - start_idx + 1), hi = min(S - 1, j + end_idx - 1);
        for (int q = lo; q <= hi; ++q) acc += sm[q];
        out[row * S + j] = from_f32<T>(acc);
    }
}

// ---------------------------------------------------------------------------- exclusive_cumprod
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
exclusive_cumprod_kernel(const T* __restrict__ x, T* __restrict__ out, int S, float eps, int inclusive,
                         unsigned* status) {
    extern __shared__ float sm[];
    float* a = sm;
    float* scratch = sm + S + 1;
    const size_t row = blockIdx.x;
    unsigned bits = 0u;
    // exclusive: a[0] = log(1 + eps), a[j+1] = log(x_j + eps): cumsum, exp, drop the last
    // inclusive (safe_cumprod): a[j] = log(x_j + eps)
    const int shift = inclusive ? 0 : 1;
    for (int j = threadIdx.x; j < S + shift; j += kRowThreads) {
        const float v = (j < shift) ? 1.0f : to_f32<T>(x[row * S + j - shift]);
        const float t = v + eps;
        if (t < 0.f) bits |= SIMULST_ST_NEGPROD;
        a[j] = logf(t);
    }
    __syncthreads();
    block_scan_smem<false>(a, S + shift, scratch);
    for (int j = threadIdx.x; j < S; j += kRowThreads) out[row * S + j] = from_f32<T>(expf(a[j]));
    flag_status(status, bits);
}

// Autograd of exclusive_cumprod / safe_cumprod (the reference's versions are differentiable
// compositions log -> cumsum -> exp, functions.py:20-66): with y the forward output,
//   d y_j / d x_k = y_j / (x_k + eps)  for j >= k + shift   (shift = 1 exclusive, 0 inclusive)
// so grad_x_k = suffix_{j >= k+shift}(g_j * y_j) / (x_k + eps).
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
cumprod_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ g,
                   T* __restrict__ gx, int S, float eps, int inclusive) {
    extern __shared__ float sm[];
    float* a = sm;
    float* scratch = sm + S + 1;
    const size_t row = blockIdx.x;
    for (int j = threadIdx.x; j < S; j += kRowThreads) a[j] = to_f32<T>(g[row * S + j]) * to_f32<T>(y[row * S + j]);
    if (threadIdx.x == 0) a[S] = 0.f;
    __syncthreads();
    block_scan_smem<true>(a, S + 1, scratch);
    const int shift = inclusive ? 0 : 1;
    for (int k = threadIdx.x; k < S; k += kRowThreads)
        gx[row * S + k] = from_f32<T>(a[k + shift] / (to_f32<T>(x[row * S + k]) + eps));
}

// ---------------------------------------------------------------------------- p_choose
template <typename T>
__global__ void p_choose_kernel(const T* __restrict__ energy, const T* __restrict__ noise,
                                T* __restrict__ out, long long numel) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
        float v = to_f32<T>(energy[i]);
        if (noise) v = to_f32<T>(from_f32<T>(v + to_f32<T>(noise[i])));     // add rounds to T first
        out[i] = from_f32<T>(1.0f / (1.0f + expf(-v)));
    }
}

template <typename F>
static int dispatch1(int dtype, F&& f) {
    switch (dtype) {
        case SIMULST_F32: return f(float{});
        case SIMULST_BF16: return f(__nv_bfloat16{});
        case SIMULST_F16: return f(__half{});
    }
    return SIMULST_E_ARG;
}

template <bool BWD>
static int launch_soft(const void* alpha, int a_dtype, const void* energy, int e_dtype,
                       const uint8_t* mask, void* beta, const void* g_beta, void* g_alpha, void* g_energy,
                       int N, int T, int S, float eps, int chunk, unsigned flags, unsigned* status,
                       cudaStream_t st) {
    if (!valid_dtype(a_dtype) || !valid_dtype(e_dtype)) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
    if ((long long)N * T == 0 || S == 0) return SIMULST_OK;
    const float fill = (flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const size_t smem = ((size_t)6 * S + kRowThreads + 16) * sizeof(float);
    if (smem > 227 * 1024) return SIMULST_E_SHAPE;      // stand-alone entry point: S <= ~9600
    return dispatch1(a_dtype, [&](auto ta) {
        using TA = decltype(ta);
        return dispatch1(e_dtype, [&](auto te) {
            using TE = decltype(te);
            auto kern = soft_attention_kernel<TA, TE, BWD>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return (int)SIMULST_E_SHAPE;
            }
            kern<<<(unsigned)((long long)N * T), kRowThreads, smem, st>>>(
                (const TA*)alpha, (const TE*)energy, mask, (TA*)beta, (const TA*)g_beta, (TA*)g_alpha,
                (TE*)g_energy, T, S, eps, chunk, fill, status);
            return check_launch();
        });
    });
}

}  // namespace simulst

using namespace simulst;

extern "C" {

int simulst_soft_attention_fwd(const void* alpha, int a_dtype, const void* soft_energy, int e_dtype,
                               const uint8_t* padding_mask, void* beta, int N, int T, int S, float eps,
                               int chunk_size, unsigned flags, unsigned* status, void* stream) {
    if (!alpha || !soft_energy || !beta || chunk_size < 0) return SIMULST_E_ARG;
    return launch_soft<false>(alpha, a_dtype, soft_energy, e_dtype, padding_mask, beta, nullptr, nullptr,
                              nullptr, N, T, S, eps, chunk_size, flags, status, (cudaStream_t)stream);
}

int simulst_soft_attention_bwd(const void* alpha, int a_dtype, const void* soft_energy, int e_dtype,
                               const uint8_t* padding_mask, const void* grad_beta, void* grad_alpha,
                               void* grad_energy, int N, int T, int S, float eps, int chunk_size,
                               unsigned flags, void* stream) {
    if (!alpha || !soft_energy || !grad_beta || !grad_alpha || !grad_energy || chunk_size < 0) return SIMULST_E_ARG;
    return launch_soft<true>(alpha, a_dtype, soft_energy, e_dtype, padding_mask, nullptr, grad_beta,
                             grad_alpha, grad_energy, N, T, S, eps, chunk_size, flags, nullptr,
                             (cudaStream_t)stream);
}

int simulst_mass_preservation_fwd(float* alpha, const uint8_t* padding_mask, float* side, int N, int T,
                                  int S, unsigned flags, unsigned* status, void* stream) {
    if (!alpha) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0) return SIMULST_E_SHAPE;
    if ((long long)N * T == 0 || S == 0) return SIMULST_OK;
    mass_preservation_fwd_kernel<<<(unsigned)((long long)N * T), kRowThreads, 0, (cudaStream_t)stream>>>(
        alpha, padding_mask, side, T, S, flags, status);
    return check_launch();
}

int simulst_mass_preservation_bwd(const float* grad_in, const uint8_t* padding_mask, const float* side,
                                  float* grad_out, int N, int T, int S, unsigned flags, void* stream) {
    if (!grad_in || !side || !grad_out) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0) return SIMULST_E_SHAPE;
    if ((long long)N * T == 0 || S == 0) return SIMULST_OK;
    mass_preservation_bwd_kernel<<<(unsigned)((long long)N * T), kRowThreads, 0, (cudaStream_t)stream>>>(
        grad_in, padding_mask, side, grad_out, T, S, flags);
    return check_launch();
}

int simulst_moving_sum(const void* x, void* out, int dtype, long long rows, int S, int start_idx,
                       int end_idx, void* stream) {
    if (!x || !out || start_idx <= 0 || end_idx <= 0 || !valid_dtype(dtype)) return SIMULST_E_ARG;
    if (rows < 0 || S < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (rows == 0 || S == 0) return SIMULST_OK;
    const size_t smem = (size_t)S * sizeof(float);
    return dispatch1(dtype, [&](auto t) {
        using T = decltype(t);
        auto kern = moving_sum_kernel<T>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<(unsigned)rows, kRowThreads, smem, (cudaStream_t)stream>>>((const T*)x, (T*)out, S, start_idx, end_idx);
        return check_launch();
    });
}

int simulst_exclusive_cumprod(const void* x, void* out, int dtype, long long rows, int S, float eps,
                              int inclusive, unsigned* status, void* stream) {
    if (!x || !out || !valid_dtype(dtype)) return SIMULST_E_ARG;
    if (rows < 0 || S < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (rows == 0 || S == 0) return SIMULST_OK;
    const size_t smem = ((size_t)S + 1 + kRowThreads + 16) * sizeof(float);
    return dispatch1(dtype, [&](auto t) {
        using T = decltype(t);
        auto kern = exclusive_cumprod_kernel<T>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<(unsigned)rows, kRowThreads, smem, (cudaStream_t)stream>>>((const T*)x, (T*)out, S, eps,
                                                                          inclusive, status);
        return check_launch();
    });
}

int simulst_cumprod_bwd(const void* x, const void* y, const void* grad_y, void* grad_x, int dtype, long long rows,
                        int S, float eps, int inclusive, void* stream) {
    if (!x || !y || !grad_y || !grad_x || !valid_dtype(dtype)) return SIMULST_E_ARG;
    if (rows < 0 || S < 0 || S > 48000) return SIMULST_E_SHAPE;
    if (rows == 0 || S == 0) return SIMULST_OK;
    const size_t smem = ((size_t)S + 1 + kRowThreads + 16) * sizeof(float);
    return dispatch1(dtype, [&](auto t) {
        using T = decltype(t);
        auto kern = cumprod_bwd_kernel<T>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return (int)SIMULST_E_SHAPE;
        }
        kern<<<(unsigned)rows, kRowThreads, smem, (cudaStream_t)stream>>>((const T*)x, (const T*)y, (const T*)grad_y,
                                                                          (T*)grad_x, S, eps, inclusive);
        return check_launch();
    });
}

int simulst_p_choose(const void* energy, const void* noise, void* out, int dtype, long long numel,
                     void* stream) {
    if (!energy || !out || !valid_dtype(dtype)) return SIMULST_E_ARG;
    if (numel < 0) return SIMULST_E_SHAPE;
    if (numel == 0) return SIMULST_OK;
    const int threads = 256;
    long long blocks = (numel + threads - 1) / threads;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    return dispatch1(dtype, [&](auto t) {
        using T = decltype(t);
        p_choose_kernel<T><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            (const T*)energy, (const T*)noise, (T*)out, numel);
        return check_launch();
    });
}

}  // extern "C"
