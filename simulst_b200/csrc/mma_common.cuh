// Pieces shared by the MMA training forward and backward kernels.
#pragma once

#include "common.cuh"

namespace simulst {

constexpr int kModeHard = 0;    // alpha only
constexpr int kModeSoftIL = 1;  // + infinite-lookback soft attention
constexpr int kModeSoftCk = 2;  // + chunkwise soft attention (moving_sum windows)

constexpr int kFwdStages = 3;  // depth of the forward row-staging ring
constexpr int kXSlots = 4;      // values exchanged per block-wide barrier (max)
constexpr int kXStride = 32;    // one float per warp per slot (<= 32 warps)

struct MmaParams {
    const void* p;          // [N,T,S]
    const void* e;          // [N,T,S] or null
    const uint8_t* mask;    // [N,S] or null
    float* alpha;           // fwd: out; bwd: saved forward output
    float* beta;            // fwd: out
    float* side;            // [N,T,2]
    const float* g_alpha;   // bwd
    const float* g_beta;    // bwd
    void* g_p;              // bwd out
    void* g_e;              // bwd out
    float* delays;          // fwd: optional [N,T] expected delays sum_j (j+1)*alpha'_ij
    const float* g_delays;  // bwd: optional [N,T] gradient w.r.t. the expected delays
    int N, T, S;
    float eps;
    int chunk;
    unsigned flags;
    unsigned* status;
    int tma;                // 1: rows staged by TMA bulk copies
    int vec_out;            // 1: fp32 rows may be accessed as float4
    int pipe;               // 1: software-pipelined kernels (hard / infinite lookback, needs tma)
    int fast;               // 1: dense fast-path backward kernel when the row qualifies
    int tma_shift;          // 1 (generic kernels only): rows are not 16-byte multiples; a bulk copy fetches the
                            // 16-byte aligned superset of a row and the consumer reads at the row's offset in it
    int row_filter;         // masked calls only: 0 all rows, 1 only rows whose mask is a right-padding mask
                            // (j >= len), 2 only the other rows -- the two passes of a masked call
    // Row pitches in ELEMENTS of each [N,T,S] tensor (the batch stride is T * pitch); S when dense.
    int ld_p, ld_e, ld_alpha, ld_beta, ld_ga, ld_gb, ld_gp, ld_ge;
    int cluster;            // 1: forward of a long row on a thread-block cluster (few rows: mma_fwd_cluster.cuh)
    int cluster_cl, cluster_threads;    // forced cluster shape (development knob), 0 = automatic
    int pitched;            // 1: some row pitch differs from S (only the SHIFT instantiations of the dense kernels
                            // and the generic kernels honour pitches; the other dense instantiations index with S)
    int shift;              // 1: dense kernels may take the row although its inputs are not 16-byte multiples
                            // and / or S is not a multiple of the per-thread element count (SHIFT
                            // instantiations: aligned-superset bulk copies, reads at the row's byte offset,
                            // a live length per row); needs 16-byte pitched OUTPUT rows of >= roundup(S, VPT)
};

// Block-wide: is this row's padding mask of the form (j >= len), i.e. no live column after a
// padded one?  One barrier; the result is uniform across the CTA.
__device__ __forceinline__ bool mask_is_right_padded(const uint8_t* __restrict__ mrow, int S) {
    bool bad = false;
    for (int j = threadIdx.x; j + 1 < S; j += blockDim.x) bad = bad || (mrow[j] != 0 && mrow[j + 1] == 0);
    return __syncthreads_or(bad ? 1 : 0) == 0;
}
// Row filter of a masked call split into a dense pass (right-padded rows) and a general pass
// (the other rows): true when this CTA's row belongs to the other pass.
__device__ __forceinline__ bool row_filtered_out(const MmaParams& prm, int n) {
    if (prm.mask == nullptr || prm.row_filter == 0) return false;
    const bool rp = mask_is_right_padded(prm.mask + (size_t)n * prm.S, prm.S);
    return (prm.row_filter == 1) != rp;
}

// Double-buffered per-warp exchange area in shared memory.
struct Xchg {
    float* base;
    int turn;
    __device__ __forceinline__ Xchg(float* b) : base(b), turn(0) {}
    __device__ __forceinline__ float* slot(int s) const {
        return base + ((turn & 1) * kXSlots + s) * kXStride;
    }
    __device__ __forceinline__ void flip() { ++turn; }
};

// Elements per vector access of a thread's VPT consecutive elements: the largest power of two
// that divides VPT and fits 16 bytes (VPT = 12 with 2-byte elements -> 4, i.e. 8-byte accesses:
// the thread's base offset 24*tid bytes is only 8-byte aligned).
template <typename T, int VPT>
__host__ __device__ constexpr int row_pack() {
    int pk = 16 / (int)sizeof(T);
    while (pk > 1 && (VPT % pk) != 0) pk /= 2;
    return pk;
}

// Read VPT consecutive elements of a staged row from shared memory in <=16-byte packs.
template <typename T, int VPT>
__device__ __forceinline__ void lds_row(const T* __restrict__ row, int j0, float (&out)[VPT]) {
    constexpr int PK = row_pack<T, VPT>();
#pragma unroll
    for (int q = 0; q < VPT / PK; ++q) {
        Pack<T, PK> pk = *reinterpret_cast<const Pack<T, PK>*>(row + j0 + q * PK);
#pragma unroll
        for (int k = 0; k < PK; ++k) out[q * PK + k] = to_f32<T>(pk.v[k]);
    }
}

// Rows that are not 16-byte multiples (generic kernels, prm.tma_shift): the bulk copy covers the
// 16-byte aligned superset [row & ~15, roundup16(row + bytes)) -- every 16-byte granule that holds a
// valid byte of an allocation is mapped, so the over-read never faults -- and the row starts
// `shift` elements into the staged copy.
__device__ __forceinline__ const void* tma_span(const void* row, unsigned bytes, unsigned& span_bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(row), a0 = a & ~static_cast<uintptr_t>(15);
    span_bytes = (static_cast<unsigned>(a - a0) + bytes + 15u) & ~15u;
    return reinterpret_cast<const void*>(a0);
}
template <typename T>
__device__ __forceinline__ int row_shift(const T* row) {
    return static_cast<int>((reinterpret_cast<uintptr_t>(row) & 15u) / sizeof(T));
}
// lds_row of a staged row that starts `shift` elements into its slot (scalar reads when shift != 0)
template <typename T, int VPT>
__device__ __forceinline__ void lds_row_shift(const T* __restrict__ row, int shift, int j0, float (&out)[VPT]) {
    if (shift == 0) {
        lds_row<T, VPT>(row, j0, out);
    } else {
#pragma unroll
        for (int k = 0; k < VPT; ++k) out[k] = to_f32<T>(row[shift + j0 + k]);
    }
}

// Dense SHIFT kernels: VPT consecutive elements of a staged row that starts `sh` BYTES into its ring
// slot (sh = the row's global address & 15; a multiple of the element size), as float2 pairs.
// The thread reads the ALIGNED 16-byte (8-byte when its span is not a 16-byte multiple) units that cover
// its elements -- one unit more than an aligned row needs, conflict-free like the aligned read; word
// loads at the shifted address would serialise 4- to 8-fold on the banks -- and realigns in registers
// with selects on CTA-uniform conditions: by 8 bytes, by 4 bytes, and for 16-bit rows that start on an
// odd element one byte permute per pair.  (A switch over the word shift that renames registers compiled
// to ~40 instructions per row; this is ~16.)
template <typename T, int VPT>
__device__ __forceinline__ void lds_row2_sh(const void* __restrict__ slot, unsigned sh, int j0, int slot_bytes,
                                            float2 (&out)[VPT / 2]) {
    constexpr int B = VPT * (int)sizeof(T);             // bytes per thread: a multiple of 8
    constexpr int U = (B % 16 == 0) ? 16 : 8;           // aligned unit
    constexpr int NWORDS = B / 4 + U / 4;               // words fetched
    // (clamped to the slot: only threads beyond the row's end can reach it, and everything they can see
    //  there is neutral)
    const int off = min(j0 * (int)sizeof(T) + (int)(sh & ~(unsigned)(U - 1)), slot_bytes - NWORDS * 4);
    const unsigned char* b = reinterpret_cast<const unsigned char*>(slot) + off;
    unsigned w[NWORDS + 1];
    if constexpr (U == 16) {
#pragma unroll
        for (int q = 0; q < NWORDS / 4; ++q) {
            const uint4 t = *reinterpret_cast<const uint4*>(b + 16 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NWORDS / 2; ++q) {
            const uint2 t = *reinterpret_cast<const uint2*>(b + 8 * q);
            w[2 * q] = t.x; w[2 * q + 1] = t.y;
        }
    }
    w[NWORDS] = 0u;
    constexpr int NX = B / 4 + 1;                       // words that can matter after the whole-word shifts
    if constexpr (U == 16) {
        const bool by8 = (sh & 8u) != 0u;
#pragma unroll
        for (int i = 0; i < NX + 1; ++i) w[i] = by8 ? w[(i + 2 <= NWORDS) ? i + 2 : NWORDS] : w[i];
    }
    {
        const bool by4 = (sh & 4u) != 0u;
#pragma unroll
        for (int i = 0; i < NX; ++i) w[i] = by4 ? w[i + 1] : w[i];
    }
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < VPT / 2; ++q) out[q] = make_float2(__uint_as_float(w[2 * q]), __uint_as_float(w[2 * q + 1]));
    } else {
        const unsigned sel = (sh & 2u) ? 0x5432u : 0x3210u;
#pragma unroll
        for (int q = 0; q < VPT / 2; ++q) {
            const unsigned v = __byte_perm(w[q], w[q + 1], sel);
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                out[q] = make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
            } else {
                out[q] = __half22float2(*reinterpret_cast<const __half2*>(&v));
            }
        }
    }
}
// MASKED rows: neutralise the copy overhang [live_end, end of its 16-byte granule) of a staged row (byte
// offsets inside the ring slot; `pattern` holds the neutral element in its low ESZ bytes).  Unrolled,
// predicated, independent stores: a counted loop costs one branch latency per element on the one warp
// every other warp waits for.  WIDE (SHIFT rows): also the following granule, which the previous row
// of this slot may have reached (rows start at varying byte offsets), when it lies inside the slot.
template <int ESZ, bool WIDE>
__device__ __forceinline__ void fix_overhang(unsigned char* slot, unsigned live_end, unsigned pattern, unsigned slot_bytes) {
    const unsigned dirt_end = (live_end + 15u) & ~15u;
#pragma unroll
    for (int k = 0; k < 16 / ESZ - 1; ++k) {
        const unsigned o = live_end + (unsigned)(k * ESZ);
        if (o < dirt_end) {
            if constexpr (ESZ == 4) *reinterpret_cast<unsigned*>(slot + o) = pattern;
            else *reinterpret_cast<unsigned short*>(slot + o) = (unsigned short)pattern;
        }
    }
    if constexpr (WIDE) {
        if (dirt_end + 16u <= slot_bytes) {
            const unsigned w = ESZ == 4 ? pattern : (pattern | (pattern << 16));
            *reinterpret_cast<uint4*>(slot + dirt_end) = make_uint4(w, w, w, w);
        }
    }
}
template <typename T>
__host__ __device__ constexpr unsigned neg_inf_bits() {
    return std::is_same<T, float>::value ? 0xff800000u : (std::is_same<T, __nv_bfloat16>::value ? 0xff80u : 0xfc00u);
}

// Element k (a run-time index, 0 <= k < VPT) of a register-resident pair array: a select tree on the bits
// of k (VPT - 1 selects) instead of VPT compare + select pairs.
template <int VPT>
__device__ __forceinline__ float pick_el(const float2 (&a)[VPT / 2], int k) {
    float v[VPT];
#pragma unroll
    for (int q = 0; q < VPT / 2; ++q) { v[2 * q] = a[q].x; v[2 * q + 1] = a[q].y; }
#pragma unroll
    for (int s = 1; s < VPT; s <<= 1) {
        const bool hi = (k & s) != 0;
#pragma unroll
        for (int i = 0; i + s < VPT; i += 2 * s) v[i] = hi ? v[i + s] : v[i];
    }
    return v[0];
}

// byte offset of row `i` inside its aligned-superset copy: (base + i * pitch_bytes) & 15
__device__ __forceinline__ unsigned sh_of(unsigned base_lo, int i, unsigned pitch_bytes) {
    return (base_lo + (unsigned)i * pitch_bytes) & 15u;
}
__device__ __forceinline__ unsigned lo32(const void* p) { return (unsigned)reinterpret_cast<uintptr_t>(p); }

// Global fp32 row access in float4 chunks when legal, scalar otherwise.
template <int VPT, bool FULL = false>
__device__ __forceinline__ void st_row_f32(float* __restrict__ row, int j0, int S, bool vec,
                                           const float (&v)[VPT]) {
#pragma unroll
    for (int q = 0; q < VPT / 4; ++q) {
        const int j = j0 + 4 * q;
        if constexpr (FULL) {
            *reinterpret_cast<float4*>(row + j) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else if (vec) {
            if (j < S) *reinterpret_cast<float4*>(row + j) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (j + k < S) row[j + k] = v[4 * q + k];
        }
    }
}
template <int VPT>
__device__ __forceinline__ void ld_row_f32(const float* __restrict__ row, int j0, int S, bool vec,
                                           float (&v)[VPT]) {
#pragma unroll
    for (int q = 0; q < VPT / 4; ++q) {
        const int j = j0 + 4 * q;
        if (vec) {
            float4 t = (j < S) ? *reinterpret_cast<const float4*>(row + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[4 * q + k] = (j + k < S) ? row[j + k] : 0.f;
        }
    }
}
// Typed global row store (gradients in the activation dtype).
template <typename T, int VPT, bool FULL = false>
__device__ __forceinline__ void st_row_t(T* __restrict__ row, int j0, int S, bool vec,
                                         const float (&v)[VPT]) {
    constexpr int PK = row_pack<T, VPT>();
#pragma unroll
    for (int q = 0; q < VPT / PK; ++q) {
        const int j = j0 + q * PK;
        if (vec) {
            if (FULL || j < S) {
                Pack<T, PK> pk;
#pragma unroll
                for (int k = 0; k < PK; ++k) pk.v[k] = from_f32<T>(v[q * PK + k]);
                *reinterpret_cast<Pack<T, PK>*>(row + j) = pk;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PK; ++k)
                if (j + k < S) row[j + k] = from_f32<T>(v[q * PK + k]);
        }
    }
}

// Shared-memory plan of the row-staging ring, identical on host and device.
struct StagePlan {
    int n_stage;       // ring depth
    int row_bytes;     // bytes reserved per staged row (multiple of 128)
    int rows;          // rows per stage
    int win_floats;    // chunkwise window scratch (floats), 0 otherwise
    __host__ __device__ int header_bytes() const { return 128 + 2 * kXSlots * kXStride * 4 + 128; }
    __host__ __device__ size_t total() const {
        return (size_t)header_bytes() + (size_t)n_stage * rows * row_bytes + (size_t)win_floats * 4;
    }
};

}  // namespace simulst
