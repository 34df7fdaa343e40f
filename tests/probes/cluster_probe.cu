// Development probe (not a test, not part of the library): what does it cost to split one row of
// the MMA training kernels over a 2-CTA cluster?  The real kernels do, per target step, ~830
// warp-instructions of arithmetic per warp and FIVE block-wide exchanges (warp scan -> one value
// per warp in shared memory -> barrier -> combine).  This probe reproduces that skeleton with a
// dependent FMA chain standing in for the arithmetic:
//   A  one 128-thread CTA per row, exchanges through __syncthreads            (today)
//   B  one 2-CTA cluster per row, 64 threads each, exchanges through DSMEM stores + cluster barrier
//      (barrier.cluster.arrive.release / wait.acquire)
//   C  same as B, exchanges through st.async to the partner's shared memory completing on the
//      partner's mbarrier (no cluster-wide barrier)
// rows = 512 on 148 SMs, 128 steps, 4 CTAs (A) / 8 CTAs (B, C) per SM by construction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

constexpr int kSteps = 128, kExch = 5;

__device__ __forceinline__ float work(float v, int n) {
    // n dependent FMAs x 4 independent chains (ILP like the kernels' 8-element rows)
    float a = v, b = v + 1.f, c = v + 2.f, d = v + 3.f;
    for (int i = 0; i < n; ++i) { a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 0.9999f, 0.25f); c = fmaf(c, 1.0002f, 0.125f); d = fmaf(d, 0.9998f, 0.0625f); }
    return (a + b) + (c + d);
}
__device__ __forceinline__ float wsum(float v) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(128, 4) probe_a(float* out, int fmas) {
    __shared__ float x[2][kExch][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v = threadIdx.x * 1e-3f;
    for (int s = 0; s < kSteps; ++s) {
        for (int e = 0; e < kExch; ++e) {
            v = work(v, fmas);
            const float w = wsum(v);
            if (lane == 0) x[s & 1][e][warp] = w;
            __syncthreads();
            v = v * 1e-3f + (x[s & 1][e][0] + x[s & 1][e][1]) + (x[s & 1][e][2] + x[s & 1][e][3]) * 1e-6f;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 8) probe_b(float* out, int fmas) {
    __shared__ float x[2][kExch][4];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    float* remote = cl.map_shared_rank(&x[0][0][0], rank ^ 1u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v = (rank * 64 + threadIdx.x) * 1e-3f;
    for (int s = 0; s < kSteps; ++s) {
        for (int e = 0; e < kExch; ++e) {
            v = work(v, fmas);
            const float w = wsum(v);
            if (lane == 0) {
                const int idx = ((s & 1) * kExch + e) * 4 + rank * 2 + warp;
                x[s & 1][e][rank * 2 + warp] = w;
                remote[idx] = w;
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            v = v * 1e-3f + (x[s & 1][e][0] + x[s & 1][e][1]) + (x[s & 1][e][2] + x[s & 1][e][3]) * 1e-6f;
        }
    }
    out[blockIdx.x * 64 + threadIdx.x] = v;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 8) probe_c(float* out, int fmas) {
    // each CTA owns an mbarrier per (parity, exchange); the partner's two warps st.async their totals
    // into this CTA's slots, completing 8 bytes of tx on it; own warps write locally and arrive.
    __shared__ __align__(8) uint64_t bar[2][kExch];
    __shared__ float x[2][kExch][4];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int i = 0; i < 2 * kExch; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[0][0] + i)), "r"(2));   // own 2 warps arrive
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cl.sync();
    float v = (rank * 64 + threadIdx.x) * 1e-3f;
    for (int s = 0; s < kSteps; ++s) {
        const unsigned par = (s >> 1) & 1;
        for (int e = 0; e < kExch; ++e) {
            v = work(v, fmas);
            const float w = wsum(v);
            uint64_t* b = &bar[s & 1][e];
            if (lane == 0) {
                x[s & 1][e][rank * 2 + warp] = w;
                // own arrival, expecting the partner's 4 bytes for the slot that mirrors this warp
                asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(4) : "memory");
                const uint32_t rslot = mapa(s32(&x[s & 1][e][rank * 2 + warp]), rank ^ 1u);
                const uint32_t rbar = mapa(s32(b), rank ^ 1u);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                             ::"r"(rslot), "r"(__float_as_uint(w)), "r"(rbar) : "memory");
            }
            asm volatile("{\n\t.reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                         ::"r"(s32(b)), "r"(par) : "memory");
            v = v * 1e-3f + (x[s & 1][e][0] + x[s & 1][e][1]) + (x[s & 1][e][2] + x[s & 1][e][3]) * 1e-6f;
        }
    }
    cl.sync();
    out[blockIdx.x * 64 + threadIdx.x] = v;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    float* out;
    cudaMalloc(&out, 1024 * 128 * sizeof(float));
    printf("rows,fmas_per_exchange,A_syncthreads_us,B_cluster_barrier_us,C_st_async_mbarrier_us,A_444rows_us\n");
    for (int fmas : {8, 16, 24, 32, 48}) {
        const float a = time_ms([&] { probe_a<<<512, 128>>>(out, fmas); });
        const float a444 = time_ms([&] { probe_a<<<444, 128>>>(out, fmas); });
        const float b = time_ms([&] { probe_b<<<1024, 64>>>(out, fmas); });
        const float c = time_ms([&] { probe_c<<<1024, 64>>>(out, fmas); });
        printf("512,%d,%.1f,%.1f,%.1f,%.1f\n", fmas, a * 1e3f, b * 1e3f, c * 1e3f, a444 * 1e3f);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    }
    return 0;
}
