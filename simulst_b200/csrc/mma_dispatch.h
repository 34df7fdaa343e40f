// Kernel-configuration table of the MMA training kernels and the per-dtype dispatchers.
#pragma once
#include <cuda_runtime.h>

namespace simulst {

struct MmaParams;

// (threads per CTA, elements per thread); capacity = product >= S
// CTA sizes in steps of one warp between the powers of two: a 1504-frame row runs 192 x 8 = 1536
// slots instead of 256 x 8 = 2048 (a quarter of the threads idle at every barrier).
#define SIMULST_MMA_CONFIGS(X) \
    X(32, 4) X(32, 8) X(64, 8) X(96, 8) X(128, 8) X(160, 8) X(192, 8) X(224, 8) X(256, 8) X(320, 8) X(384, 8) \
    X(448, 8) X(512, 8) X(512, 12) X(512, 16) X(1024, 16)

// The generic (one scan per barrier) kernels only exist at the power-of-two CTA sizes: they are the
// fallback path, and a row that fits `threads` slots fits the next power of two as well.
constexpr int generic_threads(int threads) {
    int t = 32;
    while (t < threads) t *= 2;
    return t;
}

int mma_fwd_dispatch_f32(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);
int mma_fwd_dispatch_bf16(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);
int mma_fwd_dispatch_f16(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);
int mma_bwd_dispatch_f32(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);
int mma_bwd_dispatch_bf16(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);
int mma_bwd_dispatch_f16(const MmaParams&, int mode, int threads, int vpt, cudaStream_t);

}  // namespace simulst
