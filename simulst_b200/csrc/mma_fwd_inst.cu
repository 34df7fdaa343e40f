// Instantiations of the MMA forward kernel for ONE activation dtype (selected with
// -DSIMULST_INST_DTYPE=0|1|2 so the three dtypes compile in parallel).
#include "mma_fwd.cuh"
#include "mma_fwd_pipe.cuh"
#include "mma_fwd_cluster.cuh"
#include "mma_dispatch.h"

namespace simulst {

#if SIMULST_INST_DTYPE == 0
using InstT = float;
#define INST_NAME mma_fwd_dispatch_f32
#elif SIMULST_INST_DTYPE == 1
using InstT = __nv_bfloat16;
#define INST_NAME mma_fwd_dispatch_bf16
#else
using InstT = __half;
#define INST_NAME mma_fwd_dispatch_f16
#endif

int INST_NAME(const MmaParams& prm, int mode, int threads, int vpt, cudaStream_t stream) {
    if (prm.cluster && prm.pipe && mode != kModeSoftCk) {
        const int rc = mode == kModeHard ? launch_mma_fwd_cluster<InstT, false>(prm, stream)
                                         : launch_mma_fwd_cluster<InstT, true>(prm, stream);
        /* 1 = the call does not qualify for the cluster kernel; a launch the device refuses (no cluster
           scheduling, e.g. under MIG) falls back to one CTA per row as well */
        if (rc == SIMULST_OK) return rc;
    }
#define X(TH, VP)                                                                         \
    if (threads == TH && vpt == VP) {                                                     \
        if constexpr (TH <= kPipeMaxThreads) {                                            \
            if ((prm.tma || prm.shift) && prm.pipe && mode != kModeSoftCk) {              \
                const int rc = mode == kModeHard ? launch_mma_fwd_pipe<TH, VP, InstT, false>(prm, stream) \
                                                 : launch_mma_fwd_pipe<TH, VP, InstT, true>(prm, stream); \
                if (rc != 1) return rc;         /* 1 = row too long for the pipelined kernel */ \
            }                                                                             \
        }                                                                                 \
        switch (mode) {                                                                   \
            case kModeHard: return launch_mma_fwd<generic_threads(TH), VP, InstT, kModeHard>(prm, stream); \
            case kModeSoftIL: return launch_mma_fwd<generic_threads(TH), VP, InstT, kModeSoftIL>(prm, stream); \
            case kModeSoftCk: return launch_mma_fwd<generic_threads(TH), VP, InstT, kModeSoftCk>(prm, stream); \
        }                                                                                 \
    }
    SIMULST_MMA_CONFIGS(X)
#undef X
    return SIMULST_E_ARG;
}

}  // namespace simulst
