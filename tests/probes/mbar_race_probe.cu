// Development probe: does compute-sanitizer racecheck model mbarrier arrive (release) / try_wait
// (acquire) as an ordering edge between warps?  Warp 0 writes a shared-memory word and arrives on
// an mbarrier; warp 1 waits on that mbarrier and reads the word -- correctly synchronised by the PTX
// memory model.  If racecheck reports a hazard here, its reports on the warp-specialised kernels of
// csrc/mma_sparse.cu (same pattern) are a limitation of the tool, not races.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mbar_race_probe mbar_race_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ float word[32];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        word[lane] = 1.0f + lane;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     ::"r"(s32(&bar)) : "memory");
        out[lane] = word[lane];
    }
}

int main() {
    float* out;
    cudaMalloc(&out, 32 * sizeof(float));
    probe<<<1, 64>>>(out);
    cudaDeviceSynchronize();
    float h[32];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mbar_race_probe: out[5] = %.1f (expected 6.0)\n", h[5]);
    return 0;
}
