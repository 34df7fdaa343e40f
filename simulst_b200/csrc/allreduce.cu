// Gradient all-reduce of the training-shape step over NVSwitch multicast.
//
// The MMA path shards by utterance and has no data-path collective; the training step it sits in
// all-reduces parameter gradients once per step.  An NCCL ring/tree all-reduce that overlaps the
// next step's kernels takes several SMs' worth of CTAs away from kernels that fill every SM
// (backward 347 -> 387 us at 8 ranks).  This one is a handful of CTAs: every rank owns 1/W of the
// buffer, pulls the SUM of all ranks' copies of its slice through the switch with
// `multimem.ld_reduce` (the reduction happens in the NVSwitch, NVLS) and pushes the result to all
// ranks with `multimem.st` -- one load and one store instruction per 16 bytes, no staging, no
// per-link ring steps.  The buffer is a symmetric allocation with a multicast mapping
// (torch.distributed._symmetric_memory provides the allocation, the rendezvous and the
// cross-rank barriers that bracket this kernel; see bench.py).
#include <algorithm>

#include "common.cuh"

namespace simulst {

__global__ void __launch_bounds__(512) multimem_allreduce_f32_kernel(float* __restrict__ mc, long long begin,
                                                                     long long end) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = begin + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < end; i += stride) {
        float x, y, z, w;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
                     : "l"(mc + i)
                     : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(mc + i), "f"(x), "f"(y), "f"(z), "f"(w)
                     : "memory");
    }
}

}  // namespace simulst

using namespace simulst;

extern "C" int simulst_multimem_allreduce_f32(void* multicast_ptr, long long numel, int rank, int world,
                                              int ctas, void* stream) {
    if (multicast_ptr == nullptr || numel < 0 || world <= 0 || rank < 0 || rank >= world || ctas <= 0)
        return SIMULST_E_ARG;
    if ((reinterpret_cast<uintptr_t>(multicast_ptr) % 16) != 0 || numel % 4 != 0) return SIMULST_E_ALIGN;
    if (numel == 0) return SIMULST_OK;
    // slices in units of 4 floats (one 16-byte multimem access)
    const long long vecs = numel / 4, per = (vecs + world - 1) / world;
    const long long begin = std::min(vecs, per * rank) * 4, end = std::min(vecs, per * (rank + 1)) * 4;
    if (begin >= end) return SIMULST_OK;
    multimem_allreduce_f32_kernel<<<ctas, 512, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(multicast_ptr), begin, end);
    return check_launch();
}
