// C-ABI entry points of the MMA training path (see include/simulst_b200.h).
#include <algorithm>
#include <atomic>

#include "mma_common.cuh"
#include "mma_dispatch.h"
#include "mma_sparse.h"

namespace simulst {

// Tuning overrides (development / test knobs; every default is the measured best).  Each is one
// relaxed atomic word: a setter racing with a call on another thread changes WHICH kernel variant
// that call picks, never its results (all variants pass the same parity tests), and every entry
// point snapshots the words once at its top, so one call never mixes two settings.
static std::atomic<int> g_cfg{0};          // (threads << 8) | vpt; 0 = automatic
static std::atomic<int> g_use_tma{1};
static std::atomic<int> g_use_pipe{5};     // bit 0: pipelined forward, bit 1: pipelined backward, bit 2: dense fast-path backward

struct Config { int threads, vpt; };

static bool config_exists(int threads, int vpt) {
#define X(TH, VP) if (threads == TH && vpt == VP) return true;
    SIMULST_MMA_CONFIGS(X)
#undef X
    return false;
}

// Smallest configuration that keeps a whole source row on chip.
static Config pick_config(int S) {
    const int forced = g_cfg.load(std::memory_order_relaxed);
    if (forced != 0 && (forced >> 8) * (forced & 255) >= S) return {forced >> 8, forced & 255};
    if (S <= 128) return {32, 4};
    if (S <= 256) return {32, 8};
    if (S <= 1024) return {(S + 255) / 256 * 32, 8};            // 64, 96, 128 threads
    if (S <= 2048) return {(S + 255) / 256 * 32, 8};            // 160 .. 256 threads
    if (S <= 4096) return {(S + 511) / 512 * 64, 8};            // 320 .. 512 threads
    if (S <= 6144) return {512, 12};
    if (S <= 8192) return {512, 16};
    return {1024, 16};
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int check_device() {
    static int cached[64] = {};     // 0 unknown, 1 ok, -1 wrong arch
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SIMULST_E_ARCH;
    int& c = cached[dev & 63];
    if (c == 0) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        c = (major == 10) ? 1 : -1;
    }
    return c == 1 ? SIMULST_OK : SIMULST_E_ARCH;
}

static std::atomic<int> g_cluster_shape{0};    // (cl << 16) | threads; 0 = automatic
static std::atomic<int> g_cluster{1};          // long rows in small batches: 0 never, 1 automatic, 2 whenever the shape qualifies
static std::atomic<int> g_pooled_grid{1};      // pooled calls: pooled-grid kernels (1) or always expand + dense kernels (0)
static std::atomic<int> g_split_masked{1};     // masked calls: dense pass for right-padded rows + general pass for the rest

// A masked call is split when the dense kernels can take the right-padded rows: hard or
// infinite-lookback attention, TMA-legal rows that divide evenly among the threads, no
// left-padding semantics, and no promise flag (with the promise the dense pass runs alone).
static bool split_masked_call(const MmaParams& prm, int mode, const Config& cfg, bool dense_enabled) {
    if (!g_split_masked.load(std::memory_order_relaxed) || !dense_enabled || prm.mask == nullptr || mode == kModeSoftCk) return false;
    if (prm.flags & (SIMULST_MMA_LEFT_PADDING | SIMULST_MMA_RIGHT_PADDING)) return false;
    if (cfg.threads > 512 || cfg.vpt > 12) return false;
    if (prm.shift) return true;
    if (!prm.tma || !prm.vec_out || prm.S % cfg.vpt != 0 || prm.pitched) return false;
    return true;
}

static int mode_of(unsigned flags, int chunk) {
    if (!(flags & SIMULST_MMA_SOFT)) return kModeHard;
    return chunk > 0 ? kModeSoftCk : kModeSoftIL;
}

// ---- pooled p_choose (fixed pre-decision): stand-alone expansion / gradient gather, used when a
// shape does not qualify for the kernels that expand the row in registers.
// dense[n,t,j] = pooled[n,t,(j+1)/r - 1] if (j+1) % r == 0, pooled[n,t,Sp-1] if j == S-1, else 0
// (insert_zeros + slice + last-column assignment, modules/fixed_pre_decision.py:85-95,139-159)
template <typename T>
__global__ void pool_expand_kernel(const T* __restrict__ pooled, T* __restrict__ dense, size_t rows, int S, int Sp, int r) {
    const size_t total = rows * (size_t)S;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const size_t row = q / S;
        const int j = (int)(q - row * S);
        T v = from_f32<T>(0.f);
        if (j == S - 1) v = pooled[row * Sp + Sp - 1];
        else if ((j + 1) % r == 0) v = pooled[row * Sp + (j + 1) / r - 1];
        dense[q] = v;
    }
}
// pooled_grad[n,t,k] = dense_grad[n,t, k == Sp-1 ? S-1 : (k+1)*r - 1]
template <typename T>
__global__ void pool_gather_kernel(const T* __restrict__ dense, T* __restrict__ pooled, size_t rows, int S, int Sp, int r) {
    const size_t total = rows * (size_t)Sp;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const size_t row = q / Sp;
        const int k = (int)(q - row * Sp);
        pooled[q] = dense[row * S + (k == Sp - 1 ? S - 1 : (k + 1) * r - 1)];
    }
}
template <typename T>
static int run_pool_expand(const void* pooled, void* dense, size_t rows, int S, int Sp, int r, cudaStream_t st) {
    const size_t total = rows * (size_t)S;
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    pool_expand_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(pooled), static_cast<T*>(dense), rows, S, Sp, r);
    return check_launch();
}
template <typename T>
static int run_pool_gather(const void* dense, void* pooled, size_t rows, int S, int Sp, int r, cudaStream_t st) {
    const size_t total = rows * (size_t)Sp;
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    pool_gather_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(dense), static_cast<T*>(pooled), rows, S, Sp, r);
    return check_launch();
}

// Does a pooled call run on the pooled-grid kernels (mma_sparse.cu)?  Hard or infinite-lookback
// attention, no left-padding semantics, a padding mask only with the right-padding promise,
// TMA-legal rows (S * esize a multiple of 16 bytes) of at most 4096 frames.
static bool pooled_fused_shape(int p_dtype, int S, int ratio, int chunk_size, unsigned flags, bool has_mask) {
    if (!g_use_tma.load(std::memory_order_relaxed) || !g_pooled_grid.load(std::memory_order_relaxed)) return false;
    if ((flags & SIMULST_MMA_SOFT) && chunk_size > 0) return false;
    if (flags & SIMULST_MMA_LEFT_PADDING) return false;
    if (has_mask && !(flags & SIMULST_MMA_RIGHT_PADDING)) return false;
    if (ratio < 2 || S > 4096) return false;
    return ((size_t)S * dtype_size(p_dtype)) % 16 == 0;
}

// Workspace of the pooled-grid path, carried from the forward to the backward call:
// [a_sp N*T*Sp f32][g_sp N*T*Sp f32][a_x N*T f32][g_x4 N*T float4][mp_info N*T float4][lens N i32][xcol N i32],
// segments padded to 256 bytes.
struct PooledWs {
    size_t off_asp, off_gsp, off_ax, off_gx, off_info, off_lens, off_xcol, total;
    PooledWs(int N, int T, int Sp) {
        auto pad = [](size_t b) { return (b + 255) / 256 * 256; };
        const size_t grid = pad((size_t)N * T * Sp * 4), row = pad((size_t)N * T * 4), row4 = pad((size_t)N * T * 16),
                     per_n = pad((size_t)N * 4);
        off_asp = 0; off_gsp = grid; off_ax = 2 * grid; off_gx = off_ax + row; off_info = off_gx + row4;
        off_lens = off_info + row4; off_xcol = off_lens + per_n; total = off_xcol + per_n;
    }
    void bind(SparseParams& q, void* ws) const {
        unsigned char* b = static_cast<unsigned char*>(ws);
        q.a_sp = reinterpret_cast<float*>(b + off_asp); q.g_sp = reinterpret_cast<float*>(b + off_gsp);
        q.a_x = reinterpret_cast<float*>(b + off_ax); q.g_x4 = reinterpret_cast<float4*>(b + off_gx);
        q.mp_info = reinterpret_cast<float4*>(b + off_info);
        q.lens = reinterpret_cast<int*>(b + off_lens); q.xcol = reinterpret_cast<int*>(b + off_xcol);
    }
};

// Row pitches (elements) of the [N,T,S] tensors of a call; 0 = dense (pitch S).
struct Pitches {
    long long p = 0, e = 0, alpha = 0, beta = 0, ga = 0, gb = 0, gp = 0, ge = 0;
};
static bool pitch_ok(long long ld, int S) { return ld == 0 || (ld >= S && ld <= (1ll << 30)); }
static int pitch_of(long long ld, int S) { return ld == 0 ? S : (int)ld; }
// a tensor whose every row starts on a 16-byte boundary
static bool rows16(const void* ptr, int ld, size_t esz) { return aligned(ptr, 16) && ((size_t)ld * esz) % 16 == 0; }
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

int mma_fwd_core(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                 const uint8_t* padding_mask, float* alpha, float* beta, float* side, float* expected_delays,
                 int N, int T, int S, float eps, int chunk_size, unsigned flags, unsigned* status, void* stream,
                 int pool_ratio, void* p_dense, void* workspace, const Pitches& ld = Pitches());
int mma_bwd_core(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                 const uint8_t* padding_mask, const float* alpha, const float* side, const float* grad_alpha,
                 const float* grad_beta, const float* grad_expected_delays, void* grad_p, int gp_dtype,
                 void* grad_energy, int ge_dtype, int N, int T, int S, float eps, int chunk_size, unsigned flags,
                 void* stream, int pool_ratio, const void* p_dense, void* grad_p_dense, void* workspace,
                 const Pitches& ld = Pitches());

}  // namespace simulst

using namespace simulst;

extern "C" {

int simulst_mma_set_config(int threads, int vpt) {
    if (threads == 0 && vpt == 0) { g_cfg.store(0, std::memory_order_relaxed); return SIMULST_OK; }
    if (!config_exists(threads, vpt)) return SIMULST_E_ARG;
    g_cfg.store((threads << 8) | vpt, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_set_tma(int enable) {
    g_use_tma.store(enable ? 1 : 0, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_set_cluster_shape(int cl, int threads) {
    if (cl == 0 && threads == 0) { g_cluster_shape.store(0, std::memory_order_relaxed); return SIMULST_OK; }
    if ((cl != 2 && cl != 4 && cl != 8) || (threads != 96 && threads != 128)) return SIMULST_E_ARG;
    g_cluster_shape.store((cl << 16) | threads, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_set_cluster(int mode) {
    if (mode < 0 || mode > 2) return SIMULST_E_ARG;
    g_cluster.store(mode, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_set_mask_split(int enable) {
    g_split_masked.store(enable ? 1 : 0, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_set_pipeline(int mode) {
    if (mode < 0 || mode > 7) return SIMULST_E_ARG;
    g_use_pipe.store(mode, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_train_fwd(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                          const uint8_t* padding_mask, float* alpha, float* beta, float* side,
                          int N, int T, int S, float eps, int chunk_size, unsigned flags,
                          unsigned* status, void* stream) {
    return simulst_mma_train_fwd_delays(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, beta, side,
                                        nullptr, N, T, S, eps, chunk_size, flags, status, stream);
}

int simulst_mma_train_fwd_delays(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, float* alpha, float* beta, float* side,
                                 float* expected_delays,
                                 int N, int T, int S, float eps, int chunk_size, unsigned flags,
                                 unsigned* status, void* stream) {
    return mma_fwd_core(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, beta, side, expected_delays,
                        N, T, S, eps, chunk_size, flags, status, stream, 0, nullptr, nullptr);
}

}  // extern "C"

namespace simulst {
// pool_ratio > 0: p_choose is the pooled [N,T,ceil(S/ratio)] tensor, p_dense the optional dense output
int mma_fwd_core(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                 const uint8_t* padding_mask, float* alpha, float* beta, float* side,
                 float* expected_delays,
                 int N, int T, int S, float eps, int chunk_size, unsigned flags,
                 unsigned* status, void* stream, int pool_ratio, void* p_dense, void* workspace, const Pitches& ld) {
    const bool soft = (flags & SIMULST_MMA_SOFT) != 0u;
    // the dense alpha output is optional on the pooled-grid path (soft attention only: the caller
    // then consumes alpha through beta and the expected delays)
    const bool alpha_optional = pool_ratio > 0 && soft && workspace != nullptr && p_dense == nullptr;
    if (p_choose == nullptr || (alpha == nullptr && !alpha_optional) || !valid_dtype(p_dtype)) return SIMULST_E_ARG;
    if (soft && (soft_energy == nullptr || beta == nullptr || e_dtype != p_dtype)) return SIMULST_E_ARG;
    if (chunk_size < 0) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
    if (!pitch_ok(ld.p, S) || !pitch_ok(ld.e, S) || !pitch_ok(ld.alpha, S) || !pitch_ok(ld.beta, S)) return SIMULST_E_SHAPE;
    if (N == 0 || T == 0 || S == 0) return SIMULST_OK;
    const size_t esz = dtype_size(p_dtype);
    if (!aligned(p_choose, esz) || (soft && !aligned(soft_energy, esz)) || !aligned(alpha, 4) ||
        (soft && !aligned(beta, 4)))
        return SIMULST_E_ALIGN;
    int rc = check_device();
    if (rc != SIMULST_OK) return rc;

    MmaParams prm{};
    prm.p = p_choose; prm.e = soft ? soft_energy : nullptr; prm.mask = padding_mask;
    prm.alpha = alpha; prm.beta = soft ? beta : nullptr; prm.side = side;
    prm.delays = expected_delays;
    prm.N = N; prm.T = T; prm.S = S; prm.eps = eps; prm.chunk = chunk_size; prm.flags = flags;
    prm.status = status;
    prm.ld_p = pitch_of(ld.p, S); prm.ld_e = pitch_of(ld.e, S);
    prm.ld_alpha = pitch_of(ld.alpha, S); prm.ld_beta = pitch_of(ld.beta, S);
    prm.ld_ga = prm.ld_gb = prm.ld_gp = prm.ld_ge = S;
    const int use_pipe = g_use_pipe.load(std::memory_order_relaxed);
    const bool use_tma = g_use_tma.load(std::memory_order_relaxed) != 0;
    const bool in16 = (pool_ratio > 0 || rows16(p_choose, prm.ld_p, esz)) && (!soft || rows16(soft_energy, prm.ld_e, esz));
    prm.tma = use_tma && ((size_t)S * esz) % 16 == 0 && in16;
    prm.tma_shift = use_tma && !prm.tma;
    prm.vec_out = (alpha == nullptr || rows16(alpha, prm.ld_alpha, 4)) && (!soft || rows16(beta, prm.ld_beta, 4));
    prm.pipe = use_pipe & 1;

    Config cfg = pick_config(S);
    const int mode = mode_of(flags, chunk_size);
    // A thread-block cluster per row when the rows cannot fill the SMs on their own (at most half of them) and a
    // row is long enough to be issue bound on one SM (mma_fwd_cluster.cuh)
    {
        const int cl_mode = g_cluster.load(std::memory_order_relaxed);
        // (measured at 64 rows x 128 steps, forward: S = 2048 0.88x, 3000 1.09x, 4096 1.11x, 6000 1.69x, 8192 2.26x)
        prm.cluster = pool_ratio == 0 && padding_mask == nullptr && (cl_mode == 2 || (cl_mode == 1 && N <= 74 && S > 2560)) ? 1 : 0;
        const int shape = g_cluster_shape.load(std::memory_order_relaxed);
        prm.cluster_cl = shape >> 16;
        prm.cluster_threads = shape & 0xffff;
    }
    // Dense kernels with shifted staging (SHIFT instantiations): input rows that are not 16-byte multiples
    // and / or S not a multiple of the per-thread element count, when the OUTPUT rows are 16-byte pitched
    // with room for whole threads.  The CTA then keeps 16 spare columns.
    prm.pitched = prm.ld_p != S || prm.ld_e != S || prm.ld_alpha != S || prm.ld_beta != S;
    if (use_tma && pool_ratio == 0 && prm.pipe && mode != kModeSoftCk && alpha != nullptr && S + 16 <= 6144 &&
        (!prm.tma || S % cfg.vpt != 0 || prm.pitched) && !(flags & SIMULST_MMA_LEFT_PADDING)) {
        const Config c2 = pick_config(S + 16);
        const int need = round_up(S, c2.vpt);
        if (c2.threads <= 512 && c2.vpt <= 12 && (c2.vpt < 12 || S + 32 <= 6144) && prm.vec_out && prm.ld_alpha >= need &&
            (!soft || prm.ld_beta >= need)) {
            cfg = c2;
            prm.shift = 1;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run = [&](const MmaParams& q) {
        switch (p_dtype) {
            case SIMULST_F32: return mma_fwd_dispatch_f32(q, mode, cfg.threads, cfg.vpt, st);
            case SIMULST_BF16: return mma_fwd_dispatch_bf16(q, mode, cfg.threads, cfg.vpt, st);
            default: return mma_fwd_dispatch_f16(q, mode, cfg.threads, cfg.vpt, st);
        }
    };
    if (pool_ratio > 0) {
        const int Sp = (S + pool_ratio - 1) / pool_ratio;
        if (p_dense != nullptr && !aligned(p_dense, 16)) return SIMULST_E_ALIGN;
        const bool a16 = (!soft || (aligned(soft_energy, 16) && aligned(beta, 16))) && aligned(alpha, 16);
        if (alpha == nullptr && !(workspace != nullptr && aligned(workspace, 256) && a16 &&
                                  pooled_fused_shape(p_dtype, S, pool_ratio, chunk_size, flags, padding_mask != nullptr)))
            return SIMULST_E_ARG;           // only the pooled-grid kernels can skip the dense alpha
        if (workspace != nullptr && aligned(workspace, 256) && a16 &&
            pooled_fused_shape(p_dtype, S, pool_ratio, chunk_size, flags, padding_mask != nullptr)) {
            SparseParams q{};
            q.pp = p_choose; q.e = prm.e; q.mask = padding_mask; q.alpha = alpha; q.beta = prm.beta;
            q.p_dense = p_dense; q.side = side; q.delays = expected_delays;
            q.N = N; q.T = T; q.S = S; q.Sp = Sp; q.r = pool_ratio; q.eps = eps; q.flags = flags; q.status = status;
            PooledWs(N, T, Sp).bind(q, workspace);
            return mma_sparse_run(q, p_dtype, false, st);
        }
        // the shape does not qualify: expand the row once, then the dense path
        if (p_dense == nullptr) return SIMULST_E_ARG;
        const size_t rows = (size_t)N * T;
        const int xrc = p_dtype == SIMULST_F32 ? run_pool_expand<float>(p_choose, p_dense, rows, S, Sp, pool_ratio, st)
                      : p_dtype == SIMULST_BF16 ? run_pool_expand<__nv_bfloat16>(p_choose, p_dense, rows, S, Sp, pool_ratio, st)
                                                : run_pool_expand<__half>(p_choose, p_dense, rows, S, Sp, pool_ratio, st);
        if (xrc != SIMULST_OK) return xrc;
        prm.p = p_dense;
        prm.tma = prm.tma && aligned(p_dense, 16);
        prm.tma_shift = g_use_tma.load(std::memory_order_relaxed) && !prm.tma;
    }
    if (prm.mask != nullptr && prm.shift && !(flags & SIMULST_MMA_RIGHT_PADDING) &&
        !split_masked_call(prm, mode, cfg, prm.pipe != 0))
        prm.shift = 0;      // an arbitrary mask in a single pass: generic kernels
    if (split_masked_call(prm, mode, cfg, prm.pipe != 0)) {
        // pass 1: rows whose mask is a right-padding mask, through the dense kernels;
        // pass 2: every other row, element-by-element mask handling.  Each CTA decides from its
        // own row's mask which pass it belongs to, so no host read and no extra buffer.
        MmaParams a = prm;
        a.flags |= SIMULST_MMA_RIGHT_PADDING;
        a.row_filter = 1;
        const int rc = run(a);
        if (rc != SIMULST_OK) return rc;
        MmaParams b = prm;
        b.row_filter = 2;
        return run(b);
    }
    return run(prm);
}

}  // namespace simulst

extern "C" {

int simulst_mma_train_bwd(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                          const uint8_t* padding_mask, const float* alpha, const float* side,
                          const float* grad_alpha, const float* grad_beta,
                          void* grad_p, int gp_dtype, void* grad_energy, int ge_dtype,
                          int N, int T, int S, float eps, int chunk_size, unsigned flags,
                          void* stream) {
    return simulst_mma_train_bwd_delays(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, side,
                                        grad_alpha, grad_beta, nullptr, grad_p, gp_dtype, grad_energy, ge_dtype,
                                        N, T, S, eps, chunk_size, flags, stream);
}

int simulst_mma_train_bwd_delays(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, const float* alpha, const float* side,
                                 const float* grad_alpha, const float* grad_beta,
                                 const float* grad_expected_delays,
                                 void* grad_p, int gp_dtype, void* grad_energy, int ge_dtype,
                                 int N, int T, int S, float eps, int chunk_size, unsigned flags,
                                 void* stream) {
    return mma_bwd_core(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, side, grad_alpha, grad_beta,
                        grad_expected_delays, grad_p, gp_dtype, grad_energy, ge_dtype, N, T, S, eps, chunk_size,
                        flags, stream, 0, nullptr, nullptr, nullptr);
}

}  // extern "C"

namespace simulst {
// pool_ratio > 0: p_choose / grad_p are the pooled [N,T,ceil(S/ratio)] tensors; p_dense (the
// forward's dense expansion) and grad_p_dense (workspace) serve shapes that do not qualify for
// the register-expansion kernel
int mma_bwd_core(const void* p_choose, int p_dtype, const void* soft_energy, int e_dtype,
                 const uint8_t* padding_mask, const float* alpha, const float* side,
                 const float* grad_alpha, const float* grad_beta,
                 const float* grad_expected_delays,
                 void* grad_p, int gp_dtype, void* grad_energy, int ge_dtype,
                 int N, int T, int S, float eps, int chunk_size, unsigned flags,
                 void* stream, int pool_ratio, const void* p_dense, void* grad_p_dense, void* workspace,
                 const Pitches& ld) {
    const bool soft = (flags & SIMULST_MMA_SOFT) != 0u;
    const bool mp = (flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    if (p_choose == nullptr || (alpha == nullptr && !(pool_ratio > 0 && workspace != nullptr)) || grad_p == nullptr ||
        !valid_dtype(p_dtype))
        return SIMULST_E_ARG;
    if (gp_dtype != p_dtype) return SIMULST_E_ARG;
    if (soft && (soft_energy == nullptr || grad_energy == nullptr || e_dtype != p_dtype || ge_dtype != p_dtype))
        return SIMULST_E_ARG;
    if (!soft && grad_beta != nullptr) return SIMULST_E_ARG;
    if (mp && side == nullptr) return SIMULST_E_ARG;
    if (chunk_size < 0) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
    if (!pitch_ok(ld.p, S) || !pitch_ok(ld.e, S) || !pitch_ok(ld.alpha, S) || !pitch_ok(ld.ga, S) ||
        !pitch_ok(ld.gb, S) || !pitch_ok(ld.gp, S) || !pitch_ok(ld.ge, S))
        return SIMULST_E_SHAPE;
    if (N == 0 || T == 0 || S == 0) return SIMULST_OK;
    const size_t esz = dtype_size(p_dtype);
    if (!aligned(p_choose, esz) || !aligned(grad_p, esz) || !aligned(alpha, 4)) return SIMULST_E_ALIGN;
    int rc = check_device();
    if (rc != SIMULST_OK) return rc;

    MmaParams prm{};
    prm.p = p_choose; prm.e = soft ? soft_energy : nullptr; prm.mask = padding_mask;
    prm.alpha = const_cast<float*>(alpha); prm.side = const_cast<float*>(side);
    prm.g_alpha = grad_alpha; prm.g_beta = soft ? grad_beta : nullptr;
    prm.g_delays = grad_expected_delays;
    prm.g_p = grad_p; prm.g_e = soft ? grad_energy : nullptr;
    prm.N = N; prm.T = T; prm.S = S; prm.eps = eps; prm.chunk = chunk_size; prm.flags = flags;
    prm.status = nullptr;
    prm.ld_p = pitch_of(ld.p, S); prm.ld_e = pitch_of(ld.e, S); prm.ld_alpha = pitch_of(ld.alpha, S);
    prm.ld_beta = S;
    prm.ld_ga = pitch_of(ld.ga, S); prm.ld_gb = pitch_of(ld.gb, S);
    prm.ld_gp = pitch_of(ld.gp, S); prm.ld_ge = pitch_of(ld.ge, S);
    // every input row / every output row on a 16-byte boundary
    const bool in16 = (pool_ratio > 0 || rows16(p_choose, prm.ld_p, esz)) && (!soft || rows16(soft_energy, prm.ld_e, esz)) &&
                      (alpha == nullptr || rows16(alpha, prm.ld_alpha, 4)) &&
                      (grad_alpha == nullptr || rows16(grad_alpha, prm.ld_ga, 4)) &&
                      (grad_beta == nullptr || rows16(grad_beta, prm.ld_gb, 4));
    const bool out16 = (pool_ratio > 0 || rows16(grad_p, prm.ld_gp, esz)) && (!soft || rows16(grad_energy, prm.ld_ge, esz));
    const int use_pipe = g_use_pipe.load(std::memory_order_relaxed);
    const bool use_tma = g_use_tma.load(std::memory_order_relaxed) != 0;
    prm.tma = use_tma && ((size_t)S * esz) % 16 == 0 && (S % 4 == 0) && in16 && out16;
    prm.tma_shift = use_tma && !prm.tma;
    // generic kernel: 16-byte accesses to the saved alpha rows (global loads) and the gradient rows (stores)
    prm.vec_out = out16 && (alpha == nullptr || rows16(alpha, prm.ld_alpha, 4));
    prm.pipe = (use_pipe >> 1) & 1;
    prm.fast = (use_pipe >> 2) & 1;

    Config cfg = pick_config(S);
    const int mode = mode_of(flags, chunk_size);
    // dense kernel with shifted staging: see mma_fwd_core
    prm.pitched = prm.ld_p != S || prm.ld_e != S || prm.ld_alpha != S || prm.ld_ga != S || prm.ld_gb != S ||
                  prm.ld_gp != S || prm.ld_ge != S;
    if (use_tma && pool_ratio == 0 && prm.fast && mode != kModeSoftCk && alpha != nullptr && S + 16 <= 6144 &&
        (!prm.tma || S % cfg.vpt != 0 || prm.pitched) && !(flags & SIMULST_MMA_LEFT_PADDING)) {
        const Config c2 = pick_config(S + 16);
        const int need = round_up(S, c2.vpt);
        if (c2.threads <= 512 && c2.vpt <= 12 && (c2.vpt < 12 || S + 32 <= 6144) && out16 && prm.ld_gp >= need &&
            (!soft || prm.ld_ge >= need)) {
            cfg = c2;
            prm.shift = 1;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run = [&](const MmaParams& q) {
        switch (p_dtype) {
            case SIMULST_F32: return mma_bwd_dispatch_f32(q, mode, cfg.threads, cfg.vpt, st);
            case SIMULST_BF16: return mma_bwd_dispatch_bf16(q, mode, cfg.threads, cfg.vpt, st);
            default: return mma_bwd_dispatch_f16(q, mode, cfg.threads, cfg.vpt, st);
        }
    };
    bool gather = false;
    int Sp = 0;
    if (pool_ratio > 0) {
        Sp = (S + pool_ratio - 1) / pool_ratio;
        const bool s16 = (!soft || (aligned(soft_energy, 16) && aligned(grad_energy, 16))) &&
                         (grad_alpha == nullptr || aligned(grad_alpha, 16)) &&
                         (grad_beta == nullptr || aligned(grad_beta, 16));
        if (workspace != nullptr && aligned(workspace, 256) && s16 &&
            pooled_fused_shape(p_dtype, S, pool_ratio, chunk_size, flags, padding_mask != nullptr)) {
            SparseParams q{};
            q.pp = p_choose; q.e = prm.e; q.mask = padding_mask; q.side = prm.side;
            q.g_alpha = grad_alpha; q.g_beta = prm.g_beta; q.g_delays = grad_expected_delays;
            q.g_pp = grad_p; q.g_e = prm.g_e;
            q.N = N; q.T = T; q.S = S; q.Sp = Sp; q.r = pool_ratio; q.eps = eps; q.flags = flags;
            PooledWs(N, T, Sp).bind(q, workspace);
            return mma_sparse_run(q, p_dtype, true, st);
        }
        if (p_dense == nullptr || grad_p_dense == nullptr) return SIMULST_E_ARG;
        if (!aligned(p_dense, esz) || !aligned(grad_p_dense, esz)) return SIMULST_E_ALIGN;
        prm.p = p_dense; prm.g_p = grad_p_dense;
        const bool d16 = aligned(p_dense, 16) && aligned(grad_p_dense, 16);
        prm.tma = prm.tma && d16;
        prm.tma_shift = g_use_tma.load(std::memory_order_relaxed) && !prm.tma;
        prm.vec_out = prm.vec_out && d16;
        gather = true;
    }
    auto finish = [&](int code) {
        if (code != SIMULST_OK || !gather) return code;
        const size_t rows = (size_t)N * T;
        return p_dtype == SIMULST_F32 ? run_pool_gather<float>(grad_p_dense, grad_p, rows, S, Sp, pool_ratio, st)
             : p_dtype == SIMULST_BF16 ? run_pool_gather<__nv_bfloat16>(grad_p_dense, grad_p, rows, S, Sp, pool_ratio, st)
                                       : run_pool_gather<__half>(grad_p_dense, grad_p, rows, S, Sp, pool_ratio, st);
    };
    if (prm.mask != nullptr && prm.shift && !(flags & SIMULST_MMA_RIGHT_PADDING) &&
        !split_masked_call(prm, mode, cfg, prm.fast != 0))
        prm.shift = 0;
    if (split_masked_call(prm, mode, cfg, prm.fast != 0)) {
        MmaParams a = prm;          // see simulst_mma_train_fwd_delays
        a.flags |= SIMULST_MMA_RIGHT_PADDING;
        a.row_filter = 1;
        const int rc = run(a);
        if (rc != SIMULST_OK) return rc;
        MmaParams b = prm;
        b.row_filter = 2;
        return finish(run(b));
    }
    return finish(run(prm));
}
}  // namespace simulst

extern "C" {

int simulst_mma_out_pitch(int S) {
    if (S <= 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
    const Config c1 = pick_config(S);
    if (S % 8 == 0 && S % c1.vpt == 0) return S;      // dense rows already qualify, whatever the dtype
    if (S + 16 > 6144) return round_up(S, c1.vpt > 8 ? c1.vpt : 8);
    const Config c2 = pick_config(S + 16);
    // a multiple of 8 elements (16-byte rows for every dtype) that holds whole threads of either configuration
    int ld = round_up(S, 8);
    while (ld < round_up(S, c1.vpt) || ld < round_up(S, c2.vpt) || ld % 8 != 0) ld += 4;
    return ld;
}

int simulst_mma_train_fwd_pitched(const void* p_choose, int p_dtype, long long ld_p,
                                  const void* soft_energy, int e_dtype, long long ld_e,
                                  const uint8_t* padding_mask, float* alpha, long long ld_alpha,
                                  float* beta, long long ld_beta, float* side, float* expected_delays,
                                  int N, int T, int S, float eps, int chunk_size, unsigned flags,
                                  unsigned* status, void* stream) {
    Pitches ld;
    ld.p = ld_p; ld.e = ld_e; ld.alpha = ld_alpha; ld.beta = ld_beta;
    return mma_fwd_core(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, beta, side, expected_delays,
                        N, T, S, eps, chunk_size, flags, status, stream, 0, nullptr, nullptr, ld);
}

int simulst_mma_train_bwd_pitched(const void* p_choose, int p_dtype, long long ld_p,
                                  const void* soft_energy, int e_dtype, long long ld_e,
                                  const uint8_t* padding_mask, const float* alpha, long long ld_alpha,
                                  const float* side, const float* grad_alpha, long long ld_grad_alpha,
                                  const float* grad_beta, long long ld_grad_beta,
                                  const float* grad_expected_delays,
                                  void* grad_p, int gp_dtype, long long ld_grad_p,
                                  void* grad_energy, int ge_dtype, long long ld_grad_energy,
                                  int N, int T, int S, float eps, int chunk_size, unsigned flags, void* stream) {
    Pitches ld;
    ld.p = ld_p; ld.e = ld_e; ld.alpha = ld_alpha; ld.ga = ld_grad_alpha; ld.gb = ld_grad_beta;
    ld.gp = ld_grad_p; ld.ge = ld_grad_energy;
    return mma_bwd_core(p_choose, p_dtype, soft_energy, e_dtype, padding_mask, alpha, side, grad_alpha, grad_beta,
                        grad_expected_delays, grad_p, gp_dtype, grad_energy, ge_dtype, N, T, S, eps, chunk_size,
                        flags, stream, 0, nullptr, nullptr, nullptr, ld);
}

int simulst_mma_pooled_is_fused(int p_dtype, int S, int ratio, int chunk_size, unsigned flags, int has_mask) {
    if (!valid_dtype(p_dtype) || S <= 0 || S > SIMULST_MMA_MAX_SRC || ratio < 2) return 0;
    return pooled_fused_shape(p_dtype, S, ratio, chunk_size, flags, has_mask != 0) ? 1 : 0;
}

long long simulst_mma_pooled_workspace_bytes(int N, int T, int S, int ratio) {
    if (N < 0 || T < 0 || S <= 0 || ratio < 2) return SIMULST_E_SHAPE;
    return (long long)PooledWs(N, T, (S + ratio - 1) / ratio).total;
}

int simulst_mma_set_pooled_grid(int enable) {
    g_pooled_grid.store(enable ? 1 : 0, std::memory_order_relaxed);
    return SIMULST_OK;
}

int simulst_mma_train_fwd_pooled(const void* p_pooled, int p_dtype, int ratio, const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, void* p_dense, float* alpha, float* beta, float* side,
                                 float* expected_delays, void* workspace, int N, int T, int S, float eps,
                                 int chunk_size, unsigned flags, unsigned* status, void* stream) {
    if (ratio < 2) return SIMULST_E_ARG;
    return mma_fwd_core(p_pooled, p_dtype, soft_energy, e_dtype, padding_mask, alpha, beta, side, expected_delays,
                        N, T, S, eps, chunk_size, flags, status, stream, ratio, p_dense, workspace);
}

int simulst_mma_train_bwd_pooled(const void* p_pooled, int p_dtype, int ratio, const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, const void* p_dense, const float* alpha,
                                 const float* side, const float* grad_alpha, const float* grad_beta,
                                 const float* grad_expected_delays, void* grad_p_pooled, int gp_dtype,
                                 void* grad_p_dense, void* grad_energy, int ge_dtype, void* workspace,
                                 int N, int T, int S, float eps, int chunk_size, unsigned flags, void* stream) {
    if (ratio < 2) return SIMULST_E_ARG;
    return mma_bwd_core(p_pooled, p_dtype, soft_energy, e_dtype, padding_mask, alpha, side, grad_alpha, grad_beta,
                        grad_expected_delays, grad_p_pooled, gp_dtype, grad_energy, ge_dtype, N, T, S, eps,
                        chunk_size, flags, stream, ratio, p_dense, grad_p_dense, workspace);
}

}  // extern "C"
