// C-ABI entry points of the MMA training path (see include/simulst_b200.h).
#include <atomic>

#include "mma_common.cuh"
#include "mma_dispatch.h"

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
    if (S <= 512) return {64, 8};
    if (S <= 1024) return {128, 8};
    if (S <= 2048) return {256, 8};
    if (S <= 4096) return {512, 8};
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

static std::atomic<int> g_split_masked{1};     // masked calls: dense pass for right-padded rows + general pass for the rest

// A masked call is split when the dense kernels can take the right-padded rows: hard or
// infinite-lookback attention, TMA-legal rows that divide evenly among the threads, no
// left-padding semantics, and no promise flag (with the promise the dense pass runs alone).
static bool split_masked_call(const MmaParams& prm, int mode, const Config& cfg, bool dense_enabled) {
    if (!g_split_masked.load(std::memory_order_relaxed) || !dense_enabled || prm.mask == nullptr || mode == kModeSoftCk) return false;
    if (prm.flags & (SIMULST_MMA_LEFT_PADDING | SIMULST_MMA_RIGHT_PADDING)) return false;
    if (!prm.tma || !prm.vec_out || prm.S % cfg.vpt != 0 || cfg.threads > 512 || cfg.vpt > 12) return false;
    return true;
}

static int mode_of(unsigned flags, int chunk) {
    if (!(flags & SIMULST_MMA_SOFT)) return kModeHard;
    return chunk > 0 ? kModeSoftCk : kModeSoftIL;
}

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
    const bool soft = (flags & SIMULST_MMA_SOFT) != 0u;
    if (p_choose == nullptr || alpha == nullptr || !valid_dtype(p_dtype)) return SIMULST_E_ARG;
    if (soft && (soft_energy == nullptr || beta == nullptr || e_dtype != p_dtype)) return SIMULST_E_ARG;
    if (chunk_size < 0) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
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
    const int use_pipe = g_use_pipe.load(std::memory_order_relaxed);
    prm.tma = g_use_tma.load(std::memory_order_relaxed) && ((size_t)S * esz) % 16 == 0 && aligned(p_choose, 16) &&
              (!soft || aligned(soft_energy, 16));
    prm.vec_out = (S % 4 == 0) && aligned(alpha, 16) && (!soft || aligned(beta, 16));
    prm.pipe = use_pipe & 1;

    const Config cfg = pick_config(S);
    const int mode = mode_of(flags, chunk_size);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run = [&](const MmaParams& q) {
        switch (p_dtype) {
            case SIMULST_F32: return mma_fwd_dispatch_f32(q, mode, cfg.threads, cfg.vpt, st);
            case SIMULST_BF16: return mma_fwd_dispatch_bf16(q, mode, cfg.threads, cfg.vpt, st);
            default: return mma_fwd_dispatch_f16(q, mode, cfg.threads, cfg.vpt, st);
        }
    };
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
    const bool soft = (flags & SIMULST_MMA_SOFT) != 0u;
    const bool mp = (flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    if (p_choose == nullptr || alpha == nullptr || grad_p == nullptr || !valid_dtype(p_dtype)) return SIMULST_E_ARG;
    if (gp_dtype != p_dtype) return SIMULST_E_ARG;
    if (soft && (soft_energy == nullptr || grad_energy == nullptr || e_dtype != p_dtype || ge_dtype != p_dtype))
        return SIMULST_E_ARG;
    if (!soft && grad_beta != nullptr) return SIMULST_E_ARG;
    if (mp && side == nullptr) return SIMULST_E_ARG;
    if (chunk_size < 0) return SIMULST_E_ARG;
    if (N < 0 || T < 0 || S < 0 || S > SIMULST_MMA_MAX_SRC) return SIMULST_E_SHAPE;
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
    const bool a16 = aligned(p_choose, 16) && (!soft || aligned(soft_energy, 16)) && aligned(alpha, 16) &&
                     (grad_alpha == nullptr || aligned(grad_alpha, 16)) &&
                     (grad_beta == nullptr || aligned(grad_beta, 16)) && aligned(grad_p, 16) &&
                     (!soft || aligned(grad_energy, 16));
    const int use_pipe = g_use_pipe.load(std::memory_order_relaxed);
    prm.tma = g_use_tma.load(std::memory_order_relaxed) && ((size_t)S * esz) % 16 == 0 && a16;
    prm.vec_out = ((size_t)S * esz) % 16 == 0 && (S % 4 == 0) && a16;
    prm.pipe = (use_pipe >> 1) & 1;
    prm.fast = (use_pipe >> 2) & 1;

    const Config cfg = pick_config(S);
    const int mode = mode_of(flags, chunk_size);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto run = [&](const MmaParams& q) {
        switch (p_dtype) {
            case SIMULST_F32: return mma_bwd_dispatch_f32(q, mode, cfg.threads, cfg.vpt, st);
            case SIMULST_BF16: return mma_bwd_dispatch_bf16(q, mode, cfg.threads, cfg.vpt, st);
            default: return mma_bwd_dispatch_f16(q, mode, cfg.threads, cfg.vpt, st);
        }
    };
    if (split_masked_call(prm, mode, cfg, prm.fast != 0)) {
        MmaParams a = prm;          // see simulst_mma_train_fwd_delays
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

}  // extern "C"
