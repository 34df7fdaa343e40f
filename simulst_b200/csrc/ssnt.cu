// SSNT lattice loss: the log-space twin of the MMA expected-alignment
// recurrence with the word-prediction log-probability folded in.
//
// Reference: codebase/criterion/ssnt_loss/ssnt_loss.py
//   :45-151   ssnt_loss      padded layout  [N, T, S(, V)]        (recurrence :121-127)
//   :154-271  ssnt_loss_mem  flat layout    [T_flat, S(, V)], targets concatenated over the batch
//
//   lcp[i]         = exclusive cumsum over the source axis of log(1 - p[i])
//   log_alpha[0]   = [0, neg_inf, neg_inf, ...]
//   log_alpha[i+1] = clamp(trans[i] + log_p[i] + lcp[i]
//                          + logcumsumexp(log1p(lambda) + log_alpha[i] - lcp[i]), neg_inf, 0)
//   loss[n]        = -log_alpha[n, target_len[n], source_len[n] - 1]
//
// The reference runs T rounds of ~8 eager ops over [N, S] slices plus a [N,T,S,V] gather, and its
// autograd keeps every intermediate.  Here: ONE CTA per sample walks the target axis with the
// lattice row on chip; per step an add-scan (lcp) and a log-sum-exp pair scan over the source axis;
// the backward recomputes both from the saved lattice and runs the adjoint scans (torch's
// logcumsumexp backward: sign-split reverse logcumsumexp, FunctionsManual.cpp) in the same CTA.
// Rows live in shared memory as fp32, so S <= SIMULST_SSNT_MAX_SRC; 16-bit inputs are up-cast on
// load and accumulated in fp32 (the reference accumulates lcp in the emission dtype).
//
// A separate streaming kernel (simulst_logprob_check) performs the reference's prob_check of the
// whole [*, S, V] log-prob tensor (ssnt_loss.py:29-42,80): one pass at HBM speed into the device
// status word instead of three reductions and a host sync.
#include "common.cuh"

namespace simulst {

constexpr int kSsntThreads = 256;

// ------------------------------------------------------------------ scan algebra
struct AddOp {
    using V = float;
    __device__ static V identity() { return 0.f; }
    __device__ static V combine(V a, V b) { return a + b; }
    __device__ static V shfl_up(V v, int d) { return __shfl_up_sync(kFull, v, d); }
    __device__ static V shfl_down(V v, int d) { return __shfl_down_sync(kFull, v, d); }
};
// log-sum-exp as a pair (m, s): value = m + log(s); identity (-inf, 0)
struct LseOp {
    using V = float2;
    __device__ static V identity() { return make_float2(-INFINITY, 0.f); }
    __device__ static V combine(V a, V b) {
        if (a.x >= b.x) {
            if (b.x == -INFINITY) return a;
            return make_float2(a.x, a.y + b.y * expf(b.x - a.x));
        }
        if (a.x == -INFINITY) return b;
        return make_float2(b.x, a.y * expf(a.x - b.x) + b.y);
    }
    __device__ static V shfl_up(V v, int d) {
        return make_float2(__shfl_up_sync(kFull, v.x, d), __shfl_up_sync(kFull, v.y, d));
    }
    __device__ static V shfl_down(V v, int d) {
        return make_float2(__shfl_down_sync(kFull, v.x, d), __shfl_down_sync(kFull, v.y, d));
    }
};
__device__ __forceinline__ float lse_value(float2 v) { return v.y > 0.f ? v.x + logf(v.y) : -INFINITY; }

// In-place inclusive block scan of a[0..S) in shared memory (prefix, or suffix when SUFFIX).
// Thread t owns the contiguous chunk [t*per, (t+1)*per); scratch holds kSsntThreads values.
// Every thread of the CTA calls it; contains the barriers that publish the result.
template <typename Op, bool SUFFIX>
__device__ void block_scan(typename Op::V* a, int S, typename Op::V* scratch) {
    using V = typename Op::V;
    const int tid = threadIdx.x, lane = tid & 31;
    const int per = (S + kSsntThreads - 1) / kSsntThreads;
    const int lo = min(tid * per, S), hi = min(lo + per, S);
    V run = Op::identity();
    if (!SUFFIX) {
        for (int j = lo; j < hi; ++j) { run = Op::combine(run, a[j]); a[j] = run; }
    } else {
        for (int j = hi - 1; j >= lo; --j) { run = Op::combine(run, a[j]); a[j] = run; }
    }
    scratch[tid] = run;
    __syncthreads();
    if (tid < kWarp) {
        // 256 chunk totals: lane l owns totals [8l, 8l+8)
        constexpr int Q = kSsntThreads / kWarp;
        V part[Q];
        V tot = Op::identity();
        if (!SUFFIX) {
#pragma unroll
            for (int q = 0; q < Q; ++q) { part[q] = tot; tot = Op::combine(tot, scratch[lane * Q + q]); }
            V inc = tot;
#pragma unroll
            for (int d = 1; d < kWarp; d <<= 1) {
                const V o = Op::shfl_up(inc, d);
                if (lane >= d) inc = Op::combine(o, inc);
            }
            V exc = Op::shfl_up(inc, 1);
            if (lane == 0) exc = Op::identity();
#pragma unroll
            for (int q = 0; q < Q; ++q) scratch[lane * Q + q] = Op::combine(exc, part[q]);
        } else {
#pragma unroll
            for (int q = Q - 1; q >= 0; --q) { part[q] = tot; tot = Op::combine(tot, scratch[lane * Q + q]); }
            V inc = tot;
#pragma unroll
            for (int d = 1; d < kWarp; d <<= 1) {
                const V o = Op::shfl_down(inc, d);
                if (lane + d < kWarp) inc = Op::combine(inc, o);
            }
            V exc = Op::shfl_down(inc, 1);
            if (lane == kWarp - 1) exc = Op::identity();
#pragma unroll
            for (int q = 0; q < Q; ++q) scratch[lane * Q + q] = Op::combine(exc, part[q]);
        }
    }
    __syncthreads();
    const V off = scratch[tid];
    for (int j = lo; j < hi; ++j) a[j] = Op::combine(off, a[j]);
    __syncthreads();
}

// torch.nn.functional.logsigmoid: min(0, z) - log1p(exp(-|z|))
__device__ __forceinline__ float log_sigmoid(float z) { return fminf(0.f, z) - log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.0f / (1.0f + expf(-z)); }

__host__ __device__ inline int ssnt_pitch(int S) { return (S + 3) & ~3; }

struct SsntParams {
    const void* log_probs;      // [rows, S, V]
    const int64_t* targets;     // [rows]
    const void* emit;           // [rows, S]
    const int64_t* src_len;     // [N]
    const int64_t* tgt_len;     // [N]
    const int64_t* row_off;     // [N] first input row of sample n, or null (= n*T)
    const int64_t* lat_off;     // [N] first lattice row (alpha_0) of sample n in the flat layout, or null
    float* lattice;             // padded: [N, T, S] (rows 1..T of log_alpha); flat: [T_flat + N, S] incl. alpha_0
    float* log_p;               // [rows, S] masked log p_choose (out, may be null)
    float* loss;                // [N]
    // backward
    const float* g_loss;        // [N]
    const float* g_lattice;     // same layout as lattice, or null
    const float* g_log_p;       // [rows, S] or null
    void* g_emit;               // [rows, S] emission dtype
    void* g_log_probs;          // [rows, S, V] log-prob dtype, ZERO-FILLED by the caller: the kernel writes
                                // the one gathered column per (row, frame)
    int N, T, S, V;
    int emit_is_logits;
    int flat;
    float neg_inf, fastemit_log1p;
    unsigned* status;
};

// per-sample geometry shared by forward and backward
struct SsntRow {
    long long in_row0;      // first row of emit / log_probs / targets
    long long lat_row0;     // lattice row that receives log_alpha[1]
    long long a0_row;       // flat layout: row that stores alpha_0 (-1: not stored)
    int steps;              // target steps walked for this sample
    int src, tgt;
};
__device__ __forceinline__ SsntRow ssnt_row(const SsntParams& p, int n) {
    SsntRow r;
    r.src = (int)p.src_len[n];
    r.tgt = (int)p.tgt_len[n];
    if (p.flat) {
        r.in_row0 = p.row_off[n];
        r.a0_row = p.lat_off[n];
        r.lat_row0 = r.a0_row + 1;
        r.steps = r.tgt;
    } else {
        r.in_row0 = (long long)n * p.T;
        r.a0_row = -1;
        r.lat_row0 = (long long)n * p.T;
        r.steps = p.T;
    }
    return r;
}

// log p, log(1-p) of one emission value (logits or probabilities), in fp32
__device__ __forceinline__ void emission_logs(float e, int is_logits, float& lp, float& l1) {
    if (is_logits) {
        lp = log_sigmoid(e);
        l1 = log_sigmoid(-e);
    } else {
        lp = logf(e);
        l1 = log1pf(-e);
    }
}

// One target step of the forward recurrence on shared-memory rows (also the recomputation in the
// backward).  In: a_prev[S] (log_alpha[i]).  Out: c[S] (lcp), x[S], L[S] (logcumsumexp(x)),
// y[S] (unclamped log_alpha[i+1]) and lpm[S] (masked log p).  pairs/scratch: work areas.
template <typename TE, typename TL>
__device__ void ssnt_step(const SsntParams& p, const SsntRow& r, int i, const float* a_prev, float* c, float* x,
                          float* L, float* y, float* lpm, float2* pairs, float* scratch, unsigned& bits) {
    const int S = p.S;
    const long long row = r.in_row0 + i;
    const TE* e_row = reinterpret_cast<const TE*>(p.emit) + row * S;
    const TL* lp_row = reinterpret_cast<const TL*>(p.log_probs) + row * (long long)S * p.V + p.targets[row];
    for (int j = threadIdx.x; j < S; j += kSsntThreads) {
        float lp, l1;
        const float ev = to_f32<TE>(e_row[j]);
        if (!p.emit_is_logits) bits |= prob_bits(ev);   // prob_check(emit_probs), ssnt_loss.py:84
        emission_logs(ev, p.emit_is_logits, lp, l1);
        lpm[j] = (j >= r.src) ? p.neg_inf : lp;         // source padding (ssnt_loss.py:96-98)
        c[j] = l1;
        y[j] = to_f32<TL>(lp_row[(long long)j * p.V]);  // gathered word log-prob (:113-116)
    }
    __syncthreads();
    // lcp: exclusive cumsum of log(1-p) (:22-26): inclusive scan, then shift by one
    block_scan<AddOp, false>(c, S, scratch);
    for (int j = threadIdx.x; j < S; j += kSsntThreads) x[j] = (j == 0) ? 0.f : c[j - 1];
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += kSsntThreads) {
        const float cj = x[j];
        c[j] = cj;
        const float xv = (p.fastemit_log1p + a_prev[j]) - cj;
        x[j] = xv;
        pairs[j] = make_float2(xv, 1.0f);
    }
    __syncthreads();
    block_scan<LseOp, false>(pairs, S, reinterpret_cast<float2*>(scratch));
    for (int j = threadIdx.x; j < S; j += kSsntThreads) {
        const float Lj = lse_value(pairs[j]);
        L[j] = Lj;
        y[j] = ((y[j] + lpm[j]) + c[j]) + Lj;           // (:117, :123-126)
    }
    __syncthreads();
}

__device__ __forceinline__ float clamp_logp(float v, float lo) { return fminf(fmaxf(v, lo), 0.f); }

template <typename TE, typename TL>
__global__ void __launch_bounds__(kSsntThreads)
ssnt_fwd_kernel(const SsntParams p) {
    extern __shared__ __align__(16) float sm[];
    const int S = p.S;
    const int P = ssnt_pitch(S);                        // row pitch: 16-byte aligned rows
    float* a_prev = sm;
    float* c = sm + P;
    float* x = sm + 2 * P;
    float* L = sm + 3 * P;
    float* y = sm + 4 * P;
    float* lpm = sm + 5 * P;
    float2* pairs = reinterpret_cast<float2*>(sm + 6 * P);
    float* scratch = sm + 8 * P;                        // 2 * kSsntThreads floats
    const int n = blockIdx.x;
    const SsntRow r = ssnt_row(p, n);
    bool nan_seen = false;
    unsigned bits = 0u;
    for (int j = threadIdx.x; j < S; j += kSsntThreads) a_prev[j] = (j == 0) ? 0.f : p.neg_inf;
    __syncthreads();
    if (r.a0_row >= 0)
        for (int j = threadIdx.x; j < S; j += kSsntThreads) p.lattice[r.a0_row * S + j] = a_prev[j];
    // the endpoint log_alpha[target_len, source_len - 1]; target_len == 0 reads alpha_0
    float final_val = (r.src == 1) ? 0.f : p.neg_inf;
    for (int i = 0; i < r.steps; ++i) {
        ssnt_step<TE, TL>(p, r, i, a_prev, c, x, L, y, lpm, pairs, scratch, bits);
        float* lat_row = p.lattice + (r.lat_row0 + i) * S;
        float* lp_out = p.log_p ? p.log_p + (r.in_row0 + i) * S : nullptr;
        for (int j = threadIdx.x; j < S; j += kSsntThreads) {
            const float yv = y[j];
            nan_seen = nan_seen || (yv != yv);
            const float a = clamp_logp(yv, p.neg_inf);
            a_prev[j] = a;
            lat_row[j] = a;
            if (lp_out) lp_out[j] = lpm[j];
            if (i + 1 == r.tgt && j == r.src - 1) final_val = a;
        }
        __syncthreads();
    }
    // exactly one thread saw the endpoint (or none: target_len == 0 / beyond the walked steps)
    if (r.tgt == 0) {
        if (threadIdx.x == 0) p.loss[n] = -final_val;
    } else if ((r.src - 1) % kSsntThreads == threadIdx.x && r.tgt <= r.steps) {
        p.loss[n] = -final_val;
    }
    if (nan_seen) bits |= SIMULST_ST_NAN;
    flag_status(p.status, bits);
}

// Backward: steps walked in reverse.  G = dLoss/d log_alpha[i+1] lives in shared memory.
template <typename TE, typename TL>
__global__ void __launch_bounds__(kSsntThreads)
ssnt_bwd_kernel(const SsntParams p) {
    extern __shared__ __align__(16) float sm[];
    const int S = p.S;
    const int P = ssnt_pitch(S);
    float* a_prev = sm;
    float* c = sm + P;
    float* x = sm + 2 * P;
    float* L = sm + 3 * P;
    float* y = sm + 4 * P;
    float* lpm = sm + 5 * P;
    float2* pairs = reinterpret_cast<float2*>(sm + 6 * P);
    float* G = sm + 8 * P;
    float* gy = sm + 9 * P;
    float* dx = sm + 10 * P;
    float* scratch = sm + 11 * P;
    const int n = blockIdx.x;
    const SsntRow r = ssnt_row(p, n);
    const float gl = p.g_loss ? p.g_loss[n] : 0.f;
    for (int j = threadIdx.x; j < S; j += kSsntThreads) G[j] = 0.f;
    __syncthreads();
    for (int i = r.steps - 1; i >= 0; --i) {
        // log_alpha[i]: alpha_0 or the saved lattice row i-1
        const float* prev_row = i > 0 ? p.lattice + (r.lat_row0 + i - 1) * S : nullptr;
        for (int j = threadIdx.x; j < S; j += kSsntThreads)
            a_prev[j] = prev_row ? prev_row[j] : ((j == 0) ? 0.f : p.neg_inf);
        __syncthreads();
        unsigned unused_bits = 0u;
        ssnt_step<TE, TL>(p, r, i, a_prev, c, x, L, y, lpm, pairs, scratch, unused_bits);
        // upstream of log_alpha[i+1]: carried G, the loss endpoint, an explicit lattice gradient
        const float* gl_row = p.g_lattice ? p.g_lattice + (r.lat_row0 + i) * S : nullptr;
        int any_pos = 0, any_neg = 0;
        for (int j = threadIdx.x; j < S; j += kSsntThreads) {
            float g = G[j];
            if (gl_row) g += gl_row[j];
            if (i + 1 == r.tgt && j == r.src - 1) g -= gl;
            const float yv = y[j];
            g = (yv >= p.neg_inf && yv <= 0.f) ? g : 0.f;          // clamp passes inside [neg_inf, 0]
            gy[j] = g;
            any_pos |= g > 0.f;
            any_neg |= g < 0.f;
        }
        any_pos = __syncthreads_or(any_pos);
        any_neg = __syncthreads_or(any_neg);
        // d/dx of logcumsumexp: dx_k = exp(x_k + LSE_{s>=k}(log gy+_s - L_s)) - (same with gy-)
        for (int j = threadIdx.x; j < S; j += kSsntThreads) dx[j] = 0.f;
        __syncthreads();
        for (int sign = 0; sign < 2; ++sign) {
            if (!(sign == 0 ? any_pos : any_neg)) continue;        // block-uniform
            for (int j = threadIdx.x; j < S; j += kSsntThreads) {
                const float g = sign == 0 ? gy[j] : -gy[j];
                pairs[j] = g > 0.f ? make_float2(logf(g) - L[j], 1.0f) : make_float2(-INFINITY, 0.f);
            }
            __syncthreads();
            block_scan<LseOp, true>(pairs, S, reinterpret_cast<float2*>(scratch));
            for (int j = threadIdx.x; j < S; j += kSsntThreads) {
                const float u = lse_value(pairs[j]);
                const float v = (u == -INFINITY) ? 0.f : expf(u + x[j]);
                dx[j] += sign == 0 ? v : -v;
            }
            __syncthreads();
        }
        // dc = gy - dx ; d log(1-p)_k = sum_{s>k} dc_s (adjoint of the exclusive cumsum)
        for (int j = threadIdx.x; j < S; j += kSsntThreads) c[j] = gy[j] - dx[j];
        __syncthreads();
        block_scan<AddOp, true>(c, S, scratch);
        const long long row = r.in_row0 + i;
        const TE* e_row = reinterpret_cast<const TE*>(p.emit) + row * S;
        TE* ge_row = reinterpret_cast<TE*>(p.g_emit) + row * S;
        TL* glp = p.g_log_probs ? reinterpret_cast<TL*>(p.g_log_probs) + row * (long long)S * p.V + p.targets[row] : nullptr;
        const float* gp_row = p.g_log_p ? p.g_log_p + row * S : nullptr;
        for (int j = threadIdx.x; j < S; j += kSsntThreads) {
            const float dl1 = (j + 1 < S) ? c[j + 1] : 0.f;
            float dlp = gy[j] + (gp_row ? gp_row[j] : 0.f);
            if (j >= r.src) dlp = 0.f;                              // masked_fill: constant there
            const float e = to_f32<TE>(e_row[j]);
            float ge;
            if (p.emit_is_logits) {
                const float sg = sigmoid_f(e);
                ge = dlp * (1.0f - sg) - dl1 * sg;
            } else {
                ge = dlp / e - dl1 / (1.0f - e);
            }
            ge_row[j] = from_f32<TE>(ge);
            if (glp) glp[(long long)j * p.V] = from_f32<TL>(gy[j]);
            G[j] = dx[j];                                           // flows on into log_alpha[i]
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ prob_check(log_probs, logp=True)
template <typename T>
__global__ void __launch_bounds__(256)
logprob_check_kernel(const T* __restrict__ x, long long numel, float neg_inf, unsigned* status) {
    constexpr int PK = 16 / (int)sizeof(T);
    unsigned bits = 0u;
    const long long n_vec = numel / PK;
    const Pack<T, PK>* xv = reinterpret_cast<const Pack<T, PK>*>(x);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_vec; q += (long long)gridDim.x * blockDim.x) {
        const Pack<T, PK> pk = xv[q];
#pragma unroll
        for (int k = 0; k < PK; ++k) {
            const float v = to_f32<T>(pk.v[k]);
            if (v != v) bits |= SIMULST_ST_NAN;
            if (v > 0.f || v < neg_inf) bits |= SIMULST_ST_RANGE;
        }
    }
    for (long long q = n_vec * PK + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < numel;
         q += (long long)gridDim.x * blockDim.x) {
        const float v = to_f32<T>(x[q]);
        if (v != v) bits |= SIMULST_ST_NAN;
        if (v > 0.f || v < neg_inf) bits |= SIMULST_ST_RANGE;
    }
    bits = __reduce_or_sync(kFull, bits);
    if ((threadIdx.x & 31) == 0) flag_status(status, bits);
}

template <typename F>
static int dispatch_dtype(int dtype, F&& f) {
    switch (dtype) {
        case SIMULST_F32: return f(float{});
        case SIMULST_BF16: return f(__nv_bfloat16{});
        case SIMULST_F16: return f(__half{});
        default: return SIMULST_E_ARG;
    }
}

template <typename K>
static int ssnt_launch(K kern, const SsntParams& prm, size_t smem, cudaStream_t st) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return SIMULST_E_SHAPE;
    }
    kern<<<prm.N, kSsntThreads, smem, st>>>(prm);
    return check_launch();
}

static int ssnt_check_args(const SsntParams& q, int lp_dtype, int e_dtype) {
    if (!q.log_probs || !q.targets || !q.emit || !q.src_len || !q.tgt_len || !q.lattice) return SIMULST_E_ARG;
    if (!valid_dtype(lp_dtype) || !valid_dtype(e_dtype)) return SIMULST_E_ARG;
    if (q.flat && (!q.row_off || !q.lat_off)) return SIMULST_E_ARG;
    if (q.N < 0 || q.T < 0 || q.S < 1 || q.V < 1 || q.S > SIMULST_SSNT_MAX_SRC) return SIMULST_E_SHAPE;
    return SIMULST_OK;
}

}  // namespace simulst

using namespace simulst;

extern "C" {

int simulst_ssnt_fwd(const void* log_probs, int lp_dtype, const int64_t* targets, const void* emit, int e_dtype,
                     int emit_is_logits, const int64_t* source_lengths, const int64_t* target_lengths,
                     const int64_t* row_offsets, const int64_t* lattice_offsets,
                     float* lattice, float* log_p_choose, float* loss,
                     int N, int T, int S, int V, float neg_inf, float fastemit_lambda, unsigned* status,
                     void* stream) {
    SsntParams q{};
    q.log_probs = log_probs; q.targets = targets; q.emit = emit; q.src_len = source_lengths;
    q.tgt_len = target_lengths; q.row_off = row_offsets; q.lat_off = lattice_offsets;
    q.lattice = lattice; q.log_p = log_p_choose; q.loss = loss;
    q.N = N; q.T = T; q.S = S; q.V = V; q.emit_is_logits = emit_is_logits;
    q.flat = row_offsets != nullptr;
    q.neg_inf = neg_inf; q.fastemit_log1p = log1pf(fastemit_lambda); q.status = status;
    if (loss == nullptr) return SIMULST_E_ARG;
    const int rc = ssnt_check_args(q, lp_dtype, e_dtype);
    if (rc != SIMULST_OK) return rc;
    if (N == 0) return SIMULST_OK;
    const size_t smem = ((size_t)8 * ssnt_pitch(S) + 2 * kSsntThreads + 8) * sizeof(float);
    return dispatch_dtype(e_dtype, [&](auto te) {
        return dispatch_dtype(lp_dtype, [&](auto tl) {
            return ssnt_launch(ssnt_fwd_kernel<decltype(te), decltype(tl)>, q, smem, (cudaStream_t)stream);
        });
    });
}

int simulst_ssnt_bwd(const void* log_probs, int lp_dtype, const int64_t* targets, const void* emit, int e_dtype,
                     int emit_is_logits, const int64_t* source_lengths, const int64_t* target_lengths,
                     const int64_t* row_offsets, const int64_t* lattice_offsets,
                     const float* lattice, const float* grad_loss, const float* grad_lattice,
                     const float* grad_log_p_choose, void* grad_emit, void* grad_log_probs,
                     int N, int T, int S, int V, float neg_inf, float fastemit_lambda, void* stream) {
    SsntParams q{};
    q.log_probs = log_probs; q.targets = targets; q.emit = emit; q.src_len = source_lengths;
    q.tgt_len = target_lengths; q.row_off = row_offsets; q.lat_off = lattice_offsets;
    q.lattice = const_cast<float*>(lattice);
    q.g_loss = grad_loss; q.g_lattice = grad_lattice; q.g_log_p = grad_log_p_choose;
    q.g_emit = grad_emit; q.g_log_probs = grad_log_probs;
    q.N = N; q.T = T; q.S = S; q.V = V; q.emit_is_logits = emit_is_logits;
    q.flat = row_offsets != nullptr;
    q.neg_inf = neg_inf; q.fastemit_log1p = log1pf(fastemit_lambda);
    if (grad_emit == nullptr) return SIMULST_E_ARG;
    const int rc = ssnt_check_args(q, lp_dtype, e_dtype);
    if (rc != SIMULST_OK) return rc;
    if (N == 0) return SIMULST_OK;
    const size_t smem = ((size_t)11 * ssnt_pitch(S) + 2 * kSsntThreads + 8) * sizeof(float);
    return dispatch_dtype(e_dtype, [&](auto te) {
        return dispatch_dtype(lp_dtype, [&](auto tl) {
            return ssnt_launch(ssnt_bwd_kernel<decltype(te), decltype(tl)>, q, smem, (cudaStream_t)stream);
        });
    });
}

int simulst_logprob_check(const void* log_probs, int dtype, long long numel, float neg_inf, unsigned* status,
                          void* stream) {
    if (!log_probs || !status || !valid_dtype(dtype)) return SIMULST_E_ARG;
    if (numel < 0) return SIMULST_E_SHAPE;
    if (numel == 0) return SIMULST_OK;
    if (reinterpret_cast<uintptr_t>(log_probs) % 16 != 0) return SIMULST_E_ALIGN;
    const long long want = (numel / 8 + 255) / 256;
    const int grid = (int)std::min<long long>(std::max<long long>(want, 1), 148LL * 8);
    return dispatch_dtype(dtype, [&](auto t) {
        using T = decltype(t);
        logprob_check_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)log_probs, numel, neg_inf, status);
        return check_launch();
    });
}

}  // extern "C"
