// Pooled-grid ("sparse") MMA training path: parameter block and entry point (mma_sparse.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace simulst {

struct SparseParams {
    const void* pp;         // [N,T,Sp] pooled p_choose
    const void* e;          // [N,T,S] soft energy (SIMULST_MMA_SOFT)
    const uint8_t* mask;    // [N,S] right-padding mask or null
    float* alpha;           // fwd out [N,T,S]
    float* beta;            // fwd out [N,T,S] (SOFT)
    void* p_dense;          // fwd out [N,T,S], optional
    float* side;            // [N,T,2] mass-preservation side values (fwd out, bwd in)
    float* delays;          // fwd out [N,T], optional
    // workspace (fwd writes, bwd reads): alpha on the grid, residual on an off-grid column,
    // per-row live length and off-grid column
    float* a_sp;            // [N,T,Sp]
    float* a_x;             // [N,T]
    int* lens;              // [N]   (-1: broken right-padding promise)
    int* xcol;              // [N]   (-1: none)
    // backward
    const float* g_alpha;   // [N,T,S] or null
    const float* g_beta;    // [N,T,S] or null
    const float* g_delays;  // [N,T] or null
    float* g_sp;            // workspace [N,T,Sp]: dL/d alpha' on the grid
    float4* g_x4;           // workspace [N,T] {dL/d alpha' at the off-grid residual column, -, -, -}
    float4* mp_info;        // workspace [N,T] {row sum, raw alpha at the mass-preservation column, -, -} (16-byte
                            // records so the backward can fetch step chunks by bulk copy)
    void* g_pp;             // out [N,T,Sp]
    void* g_e;              // out [N,T,S] (SOFT)
    int N, T, S, Sp, r;
    float eps;
    unsigned flags;
    unsigned* status;
};

int mma_sparse_run(const SparseParams& prm, int dtype, bool bwd, cudaStream_t st);

}  // namespace simulst
