// MMA training forward, software-pipelined: expected alignment (+ mass preservation)
// (+ infinite-lookback expected soft attention), one CTA per (batch*head) row.
//
// The only true dependency between target steps is  alpha_{i-1} -> u-prefix -> alpha_i.
// Everything else is either a function of the inputs of one step (exclusive cumprod, clamped
// divisor, P, row max, exp, D prefix) or hangs off alpha_i without feeding back (row sum,
// r-suffix, beta_i).  The kernel therefore keeps FOUR steps in flight and lets all their
// block-wide scans share ONE __syncthreads per loop iteration `it`:
//
//   stage MAXS(it+2)  row max of soft_energy                         (REDUX)
//   stage INV (it+1)  cumprod prefix-product, exp + D prefix-sum     (2 warp scans)
//   stage RECU(it)    u-prefix over alpha_{it-1}  -> alpha_it        (1 warp scan)
//   stage RECR(it-1)  r-suffix, row sum           -> beta_{it-1}, mass-preserved column
//
//   PRE : thread-local chains + warp scans of all four stages, one value per warp and scan
//         published in shared memory
//   ---- barrier (then: TMA refill of the ring slot INV just consumed)
//   POST: cross-warp combines, element-wise finish of every stage, stores
//
// so the per-step latency is one scan deep instead of four, and the four independent shuffle
// chains interleave in one instruction stream.  Rows of p_choose / soft_energy arrive through
// a ring of TMA 1-D bulk copies (UBLKCP) issued NS steps ahead by one thread; exp(E - m) + eps
// waits for its beta in a thread-private shared-memory stash (3 steps deep).
// Math / reference lines: mma_steps.cuh and mma_fwd.cuh.  Chunkwise soft attention, rows that
// TMA cannot stage (not 16-byte aligned) and rows too long for two ring stages use the generic
// kernel in mma_fwd.cuh.
#pragma once

#include "mma_common.cuh"
#include "mma_steps.cuh"

namespace simulst {

constexpr int kPipeStages = 4;          // deepest row-staging ring (shallower when rows are long)
constexpr int kPipeMaxThreads = 512;    // 1024-thread CTAs (64 registers/thread) keep the generic kernel
constexpr int kPipeSlots = 8;           // values exchanged per barrier (6 used)
constexpr int kExStash = 3;             // steps between exp(E - m) and its use in beta

struct PipePlan {
    int n_stage;       // ring depth (>= 2 with soft attention)
    int row_bytes;     // bytes reserved per staged row (multiple of 128)
    int rows;          // rows per stage (p [, energy])
    int stash_bytes;   // exp stash: kExStash * THREADS * VPT * 4, 0 without soft attention
    __host__ __device__ int header_bytes() const { return 128 + 2 * kPipeSlots * kXStride * 4 + 128; }
    __host__ __device__ size_t total() const {
        return (size_t)header_bytes() + (size_t)n_stage * rows * row_bytes + (size_t)stash_bytes;
    }
};

// Compile-time shared-memory plan of one configuration (same arithmetic as PipePlan): ring
// depth, row pitch and stash size are constants in the kernel, so ring-slot addresses cost no
// run-time multiplies.
template <int THREADS, int VPT, typename T, bool SOFT>
struct PipeStatic {
    static constexpr int kRows = SOFT ? 2 : 1;
    static constexpr int kRowBytes = ((THREADS * VPT * (int)sizeof(T)) + 127) / 128 * 128;
    static constexpr int kStashBytes = SOFT ? kExStash * THREADS * VPT * 4 : 0;
    static constexpr int kHeader = 128 + 2 * kPipeSlots * kXStride * 4 + 128;
    // keep 4 CTAs of a 128-thread configuration resident (<= 56 KB each); long rows: what fits
    static constexpr size_t kBudget = THREADS <= 128 ? 56 * 1024 : (THREADS <= 256 ? 110 * 1024 : 220 * 1024);
    static constexpr size_t total(int ns) { return (size_t)kHeader + (size_t)ns * kRows * kRowBytes + (size_t)kStashBytes; }
    static constexpr int pick() {
        int ns = kPipeStages;
        while (ns > 1 && total(ns) > kBudget) --ns;
        return ns;
    }
    static constexpr int kNS = pick();
    static constexpr bool kFits = total(kNS) <= kBudget && (!SOFT || kNS >= 2);
    static constexpr size_t kTotal = total(kNS);
};

// DELAYS: the expected-delay epilogue is compiled in (dense rows: a separate instantiation, so
// the plain kernel carries none of it; ragged / masked rows: always compiled in, run-time flag)
// RAGGED (with FULL): S < THREADS*VPT, S a multiple of VPT -- threads are wholly inside or wholly
// outside the row; outside threads initialise the ring tails to neutral values once (p = 0,
// energy = -inf; the bulk copies only write [0, S)), compute without bounds checks, skip stores.
// MASKED (with FULL and RAGGED): padding_mask is a RIGHT-padding mask (caller's promise,
// SIMULST_MMA_RIGHT_PADDING): row n is live on [0, L_n).  Only the LIVE bytes of a row are copied
// (rounded up to the 16-byte granule), every ring slot beyond them is initialised once to neutral
// values (p = 0, energy = -inf), and the thread that owns column L_n - 1 (the "fixer") overwrites the
// few columns of copy overhang behind the row's end with neutral values one step before the row is
// used -- so the step loop reads every row without per-element tests (the earlier version masked in
// registers under divergent branches: the warp the boundary falls into ran ~100 extra instructions
// per step and every other warp waited for it at the barrier, +28 % at the training shape).  alpha
// and beta come out zero beyond L_n, and mass preservation ADDS its residual at L_n - 1
// (monotonic_attention.py:186-193).
// SHIFT (with FULL, RAGGED and MASKED): rows of the INPUT tensors need not be 16-byte multiples and S
// need not be a multiple of VPT.  Every row is fetched as its 16-byte aligned superset and read at its
// byte offset inside the staged copy (lds_row2_sh); the live length is min(S, mask length) -- the mask
// is optional, without one the reference's no-mask rule applies (column S-1 REPLACED) -- and the
// thread the row ends in neutralises its columns beyond it like a masked row's boundary thread.
// The OUTPUT rows must be 16-byte pitched with room for whole threads (pitch >= roundup(S, VPT)):
// the padding columns receive zeros.  The CTA must have 16 spare columns (S + 16 <= THREADS*VPT).
template <int THREADS, int VPT, typename T, bool SOFT, bool FULL, bool DELAYS, bool RAGGED = false, bool MASKED = false,
          bool SHIFT = false>
__global__ void __launch_bounds__(THREADS, (THREADS * VPT <= 1024 ? 4 : (THREADS <= 160 ? 3 : (THREADS <= 256 ? 2 : 1))))
mma_fwd_pipe_kernel(const MmaParams prm) {
    using PS = PipeStatic<THREADS, VPT, T, SOFT>;
    constexpr int NW = THREADS / kWarp;
    constexpr int kIssuers = (SOFT && NW > 1) ? 2 : 1;     // warps that issue TMA copies
    constexpr int H = VPT / 2;
    static_assert(VPT % 4 == 0, "VPT must be a multiple of 4");

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xbuf = reinterpret_cast<float*>(smem + 128);               // [2][kPipeSlots][32]
    float* bcast = xbuf + 2 * kPipeSlots * kXStride;                  // [4] 1/D at the mass-preservation column
    unsigned char* stage0 = smem + PS::kHeader;
    float4* stash = reinterpret_cast<float4*>(stage0 + PS::kNS * PS::kRows * PS::kRowBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(kFull, tid >> 5, 0)   /* shuffle: known warp-uniform */;
    const int n = blockIdx.x;
    if (row_filtered_out(prm, n)) return;       // this row belongs to the call's other pass
    const int S = prm.S, T_len = prm.T;
    const int j0 = tid * VPT;
    const float eps = prm.eps;
    const bool mp = (prm.flags & SIMULST_MMA_MASS_PRESERVATION) != 0u;
    const float fill = (prm.flags & SIMULST_MMA_ENERGY_F16_FILL) ? -1e4f : -1e8f;
    const bool vec_out = FULL || prm.vec_out != 0;
    constexpr int NS = PS::kNS;

    // row pitches: only the SHIFT instantiation takes pitched tensors; everywhere else the pitch is S, which
    // lets the compiler share one row offset between the four tensors
    const int ld_p = SHIFT ? prm.ld_p : S, ld_e = SHIFT ? prm.ld_e : S, ld_a = SHIFT ? prm.ld_alpha : S,
              ld_b = SHIFT ? prm.ld_beta : S;
    const T* gp = reinterpret_cast<const T*>(prm.p) + (size_t)n * T_len * ld_p;
    const T* ge = SOFT ? reinterpret_cast<const T*>(prm.e) + (size_t)n * T_len * ld_e : nullptr;
    float* g_alpha = prm.alpha + (size_t)n * T_len * ld_a;
    float* g_beta = SOFT ? prm.beta + (size_t)n * T_len * ld_b : nullptr;
    // SHIFT: byte offset of a row inside its staged aligned superset, from the low address bits
    const unsigned p_lo = lo32(gp), e_lo = lo32(ge);
    const unsigned p_pitch = (unsigned)ld_p * (unsigned)sizeof(T), e_pitch = (unsigned)ld_e * (unsigned)sizeof(T);
    (void)p_lo; (void)e_lo; (void)p_pitch; (void)e_pitch;

    // ---- per-row constants: validity bits, padding mask, column rewritten by mass preservation
    unsigned in_bits = 0u, live_bits = 0u;
    int n_live = 0;
    if constexpr (!FULL) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = j0 + k;
            if (j < S) {
                in_bits |= 1u << k;
                const bool padded = prm.mask != nullptr && prm.mask[(size_t)n * S + j] != 0;
                if (!padded) { live_bits |= 1u << k; ++n_live; }
            }
        }
    }
    auto is_in = [&](int k) -> bool { return FULL ? true : ((in_bits >> k) & 1u) != 0u; };
    auto is_live = [&](int k) -> bool { return FULL ? true : ((live_bits >> k) & 1u) != 0u; };
    // mass_preservation: no mask / left padding -> REPLACE column S-1 with the residual of the
    // other columns; right padding -> ADD the residual of all columns at src_len-1.
    if constexpr (FULL && MASKED) {
        if (j0 < S) {
            if (SHIFT && prm.mask == nullptr) {
                n_live = min(VPT, S - j0);
            } else {
                const uint8_t* mrow = prm.mask + (size_t)n * S + j0;
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (!SHIFT || j0 + k < S) n_live += (mrow[k] == 0) ? 1 : 0;
            }
        }
    }
    const bool count_live = (!FULL || MASKED) && (prm.mask != nullptr || SHIFT);
    const bool mp_add = (!FULL || MASKED) && prm.mask != nullptr && !(prm.flags & SIMULST_MMA_LEFT_PADDING);
    int last = S - 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], kIssuers);
        mbar_fence_init();
    }
    const bool inside = !RAGGED || j0 < S;
    if (RAGGED && !MASKED && !inside) {
        const T ninf = from_f32<T>(-INFINITY), zero = from_f32<T>(0.f);
        for (int s = 0; s < NS; ++s) {
            T* sp = reinterpret_cast<T*>(stage0 + (s * PS::kRows) * PS::kRowBytes);
            T* se = reinterpret_cast<T*>(stage0 + (s * PS::kRows + 1) * PS::kRowBytes);
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                sp[j0 + k] = zero;
                if (SOFT) se[j0 + k] = ninf;
            }
        }
    }
    if (SHIFT ? count_live : mp_add) {
        const float cnt = warp_sum((float)n_live);
        if (lane == 0) xbuf[warp] = cnt;
    }
    __syncthreads();
    if (SHIFT ? count_live : mp_add) {
        last = (int)combine_sum<NW>(xbuf, lane) - 1;
        // (through a shuffle: the compiler then knows the value is warp-uniform and keeps the byte counts and
        //  addresses derived from it in the uniform datapath)
        if constexpr (MASKED) last = __shfl_sync(kFull, last, 0);
        __syncthreads();
    }
    // element of this thread that sits on the mass-preservation column (-1: none)
    // live columns of this thread (right-padded rows); VPT everywhere else
    const int nl = MASKED ? max(0, min(VPT, last + 1 - j0)) : VPT;
    (void)nl;
    const int L_row = last + 1;                                     // MASKED: live length of the row
    const bool fixer = MASKED && last >= j0 && last < j0 + VPT;     // owner of the last live column
    (void)L_row; (void)fixer;
    if constexpr (FULL && MASKED) {
        if (nl < VPT) {
            // neutral ring slots from this thread's columns on (the copies rewrite the live bytes)
            const T ninf = from_f32<T>(-INFINITY), zero = from_f32<T>(0.f);
            for (int s = 0; s < PS::kNS; ++s) {
                T* sp = reinterpret_cast<T*>(stage0 + (s * PS::kRows) * PS::kRowBytes);
                T* se = reinterpret_cast<T*>(stage0 + (s * PS::kRows + 1) * PS::kRowBytes);
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    sp[j0 + k] = zero;
                    if (SOFT) se[j0 + k] = ninf;
                }
            }
            fence_proxy_async();
        }
    }
    if constexpr (FULL && MASKED) {
        // the promise is checked, not trusted: a row whose mask is not (j >= len) flags the status
        // word AND has its outputs poisoned with NaN, so a broken promise cannot train on silently
        // even when nobody reads the (lazily inspected) status word
        bool ok = true;
        if (j0 < S && !(SHIFT && prm.mask == nullptr)) {
            const uint8_t* mrow = prm.mask + (size_t)n * S + j0;
#pragma unroll
            for (int k = 0; k < VPT; ++k)
                if (!SHIFT || j0 + k < S) ok = ok && ((mrow[k] == 0) == (k < nl));
        }
        if (__syncthreads_or(ok ? 0 : 1)) {
            if (tid == 0 && prm.status != nullptr) atomicOr(prm.status, SIMULST_ST_NOT_RIGHT_PADDED);
            const float qnan = __int_as_float(0x7fc00000);
            for (int i = 0; i < T_len; ++i)
                for (int j = tid; j < S; j += THREADS) {
                    g_alpha[(size_t)i * ld_a + j] = qnan;
                    if (SOFT) g_beta[(size_t)i * ld_b + j] = qnan;
                }
            return;
        }
    }
    if constexpr (FULL && MASKED) {
        // (a barrier reduction: the compiler then knows the exit is taken by the whole CTA or by nobody;
        //  a plain data-dependent return makes it treat the rest of the kernel as divergent code and give
        //  up the uniform datapath)
        if (__syncthreads_or(last < 0 ? 1 : 0)) {
            // no live column: alpha = beta = 0 for the whole row (no mass-preservation column exists)
            float2 z2[H];
#pragma unroll
            for (int q = 0; q < H; ++q) z2[q] = f2(0.f);
            if (inside)
                for (int i = 0; i < T_len; ++i) {
                    st_row2_f32<VPT, FULL>(g_alpha + (size_t)i * ld_a, j0, S, vec_out, z2);
                    if (SOFT) st_row2_f32<VPT, FULL>(g_beta + (size_t)i * ld_b, j0, S, vec_out, z2);
                }
            return;
        }
    }
    int k_last = -1;
    if (FULL && !MASKED) {
        if (RAGGED ? (j0 + VPT == S) : (tid == THREADS - 1)) k_last = VPT - 1;
    } else if (last >= j0 && last < j0 + VPT) {
        k_last = last - j0;
    }
    const bool own_last = mp && k_last >= 0;
    auto at_last = [&](int k) -> bool { return ((FULL && !MASKED) ? k == VPT - 1 : true) && k == k_last; };

    // ---- row staging ring
    const unsigned row_bytes = (unsigned)((MASKED ? L_row : S) * sizeof(T));
    auto stage_p = [&](int s) { return reinterpret_cast<T*>(stage0 + (s * PS::kRows) * PS::kRowBytes); };
    auto stage_e = [&](int s) { return reinterpret_cast<T*>(stage0 + (s * PS::kRows + 1) * PS::kRowBytes); };
    // one elected lane of warp 0 copies the p row, one of warp 1 (if there is one) the energy
    // row; each arrives on the slot's barrier with its own byte count (warp-uniform branches)
    auto issue = [&](int i, int s) {
        if (warp == 0) {
            if (elect_one()) {
                if constexpr (MASKED) {
                    unsigned np = 0u, ne = 0u;
                    const void* sp = tma_span(gp + (size_t)i * ld_p, row_bytes, np);
                    const void* se = (SOFT && kIssuers == 1) ? tma_span(ge + (size_t)i * ld_e, row_bytes, ne) : nullptr;
                    mbar_expect_tx(&bars[s], np + ne);
                    tma_load_1d(stage_p(s), sp, np, &bars[s]);
                    if (SOFT && kIssuers == 1) tma_load_1d(stage_e(s), se, ne, &bars[s]);
                } else {
                    mbar_expect_tx(&bars[s], (SOFT && kIssuers == 1) ? 2u * row_bytes : row_bytes);
                    tma_load_1d(stage_p(s), gp + (size_t)i * ld_p, row_bytes, &bars[s]);
                    if (SOFT && kIssuers == 1) tma_load_1d(stage_e(s), ge + (size_t)i * ld_e, row_bytes, &bars[s]);
                }
            }
        } else if (SOFT && kIssuers == 2 && warp == 1) {
            if (elect_one()) {
                if constexpr (MASKED) {
                    unsigned ne = 0u;
                    const void* se = tma_span(ge + (size_t)i * ld_e, row_bytes, ne);
                    mbar_expect_tx(&bars[s], ne);
                    tma_load_1d(stage_e(s), se, ne, &bars[s]);
                } else {
                    mbar_expect_tx(&bars[s], row_bytes);
                    tma_load_1d(stage_e(s), ge + (size_t)i * ld_e, row_bytes, &bars[s]);
                }
            }
        }
    };
    // SHIFT: staged row -> float2 pairs; reads are clamped to the slot (threads beyond the row read
    // neutral bytes wherever they land)
    auto load_sh = [&](const T* slot, unsigned sh, float2 (&v)[H]) {
        if (sh == 0u) {                     // CTA-uniform: this row happens to be aligned
            unsigned dummy = 0u;
            lds_row2<T, VPT, false>(slot + j0, v, dummy);
        } else {
            lds_row2_sh<T, VPT>(slot, sh, j0, THREADS * VPT * (int)sizeof(T), v);
        }
    };
    // MASKED: the fixer neutralises the copy overhang behind the row's end in a landed row
    auto fix_row = [&](T* slot, unsigned sh, unsigned pattern) {
        fix_overhang<(int)sizeof(T), SHIFT>(reinterpret_cast<unsigned char*>(slot), sh + (unsigned)L_row * (unsigned)sizeof(T),
                                            pattern, (unsigned)(THREADS * VPT * sizeof(T)));
    };
    const float eps_x = (MASKED && nl == 0) ? 0.f : eps;    // no eps from the columns of a thread beyond the row
    unsigned umax32 = 0u;                   // SHIFT: first-level prob_check on the fp32 bit patterns
    for (int i = 0; i < NS && i < T_len; ++i) issue(i, i);

    const float one_eps = 1.0f + eps;       // first element of the exclusive cumprod (functions.py:28-33)
    // state carried between iterations
    float2 a_prev[H];                       // alpha_{it-1} (before mass preservation)
    float2 rc[H], P[H], rD[H];              // invariants of the step RECU runs next
    float2 Rl[H];                           // thread-local r-suffix of the step RECR finishes next
#pragma unroll
    for (int q = 0; q < H; ++q) {
        a_prev[q] = make_float2((j0 + 2 * q == 0) ? 1.0f : 0.0f, 0.0f);
        rc[q] = P[q] = rD[q] = Rl[q] = f2(0.f);
    }
    float m_cur = 0.f;                      // row max of the step INV runs next
    float rt_prev = 0.f, rsum_prev = 0.f;   // thread totals (r-suffix, row sum) of the step RECR scans next
    float wsum_prev = 0.f;                  // thread total of (j+1)*alpha (expected delay) of that step
    const bool want_d = DELAYS && prm.delays != nullptr;
    float a_last_raw = 0.f;                 // owner thread: alpha at the mass-preservation column
    unsigned umax = 0u;                     // FULL rows: first-level prob_check
    bool bad = false;                       // ragged rows: exact per-element check
    bool nan_out = false;

    // ring slot / mbarrier parity of row it+1 (INV); MAXS reads the next slot.  The first
    // iteration is it = -2, i.e. "row -1": the last slot of the previous lap.
    int slotI = NS - 1;
    unsigned parI = 1u;
    int sb_w = kExStash - 1;                // exp-stash buffer of step it+1 (INV writes it); RECR(it-1) reads the next one
    // MASKED: this thread's output addresses as running pointers (row `it` of alpha, row `it - 1` of beta).  With
    // the row index multiplied out at every store the compiler rebuilt the 64-bit row address from %ctaid each
    // step (~24 integer multiply-adds per step in these register-tight instantiations).
    float* pa_run = MASKED ? g_alpha - 2 * (ptrdiff_t)ld_a + j0 : nullptr;
    float* pb_run = (MASKED && SOFT) ? g_beta - 3 * (ptrdiff_t)ld_b + j0 : nullptr;
    (void)pa_run; (void)pb_run;

    auto body = [&](auto steady_c, const int it) __attribute__((always_inline)) {
        constexpr bool STEADY = decltype(steady_c)::value;
        const bool doM = SOFT && (STEADY || it + 2 < T_len);
        const bool doI = STEADY || (it >= -1 && it + 1 < T_len);
        const bool doU = STEADY || (it >= 0 && it < T_len);
        const bool doR = STEADY || (it >= 1 && it - 1 < T_len);
        float* xw = xbuf + (it & 1) * (kPipeSlots * kXStride);

        // ================================================================ PRE
        // thread-local chains of every stage first, then ONE fused warp-scan section (stages that
        // are off in an edge iteration contribute identities), then the per-warp values go to
        // shared memory.
        // ---- MAXS(it+2)
        float em = -INFINITY;
        if (SOFT && doM) {
            int slotM = slotI + 1;
            unsigned parM = parI;
            if (slotM == NS) { slotM = 0; parM ^= 1u; }
            mbar_wait(&bars[slotM], parM);
            if (MASKED && fixer) {
                fix_row(stage_p(slotM), SHIFT ? sh_of(p_lo, it + 2, p_pitch) : 0u, 0u);
                fix_row(stage_e(slotM), SHIFT ? sh_of(e_lo, it + 2, e_pitch) : 0u, neg_inf_bits<T>());
            }
            if constexpr (SHIFT) {
                float2 Em[H];
                load_sh(stage_e(slotM), sh_of(e_lo, it + 2, e_pitch), Em);
                em = fmaxf(Em[0].x, Em[0].y);
#pragma unroll
                for (int q = 1; q < H; ++q) em = fmaxf(em, fmaxf(Em[q].x, Em[q].y));
            } else if constexpr (FULL && sizeof(T) == 2 && VPT % 8 == 0) {
                // packed 16-bit max, one conversion at the end
                using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
                const uint4* src = reinterpret_cast<const uint4*>(stage_e(slotM) + j0);
                T2 acc;
#pragma unroll
                for (int q = 0; q < VPT / 8; ++q) {
                    const uint4 w = src[q];
                    const T2 a = __hmax2(*reinterpret_cast<const T2*>(&w.x), *reinterpret_cast<const T2*>(&w.y));
                    const T2 b = __hmax2(*reinterpret_cast<const T2*>(&w.z), *reinterpret_cast<const T2*>(&w.w));
                    const T2 c = __hmax2(a, b);
                    acc = q == 0 ? c : __hmax2(acc, c);
                }
                em = fmaxf(to_f32<T>(acc.x), to_f32<T>(acc.y));
            } else {
                float2 Em[H];
                unsigned dummy = 0u;
                lds_row2<T, VPT, false>(stage_e(slotM) + j0, Em, dummy);
                if constexpr (!FULL) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (!is_live(k)) SIMULST_EL(Em, k) = is_in(k) ? fill : -INFINITY;
                }
                em = fmaxf(Em[0].x, Em[0].y);
#pragma unroll
                for (int q = 1; q < H; ++q) em = fmaxf(em, fmaxf(Em[q].x, Em[q].y));
            }
        }
        if (MASKED && nl == 0) em = -INFINITY;      // a thread beyond the row may have read overhang not yet fixed
        if (MASKED && !SOFT && fixer && (STEADY || it + 2 < T_len)) {
            // hard attention has no MAXS stage: the fixer alone looks two rows ahead
            int slotM = slotI + 1;
            unsigned parM = parI;
            if (slotM == NS) { slotM = 0; parM ^= 1u; }
            mbar_wait(&bars[slotM], parM);
            fix_row(stage_p(slotM), SHIFT ? sh_of(p_lo, it + 2, p_pitch) : 0u, 0u);
        }
        // ---- INV(it+1): local chains
        float2 p_n[H], cpre[H], Dl[H];
        float xinc = 1.f, einc = 0.f;
        if (doI) {
            if (!SOFT) mbar_wait(&bars[slotI], parI);      // with SOFT, MAXS waited for this row one iteration ago
            float2 E_n[H];
            if constexpr (SHIFT) {
                load_sh(stage_p(slotI), sh_of(p_lo, it + 1, p_pitch), p_n);
                if (SOFT) load_sh(stage_e(slotI), sh_of(e_lo, it + 1, e_pitch), E_n);
            } else {
                lds_row2<T, VPT, FULL>(stage_p(slotI) + j0, p_n, umax);
                if (SOFT) {
                    unsigned dummy = 0u;
                    lds_row2<T, VPT, false>(stage_e(slotI) + j0, E_n, dummy);
                }
            }
            if constexpr (SHIFT) {
                // a valid probability has the bit pattern of a float in [+0, 1]: unsigned compare
#pragma unroll
                for (int q = 0; q < H; ++q)
                    umax32 = max(umax32, max(__float_as_uint(p_n[q].x), __float_as_uint(p_n[q].y)));
            }
            if constexpr (!FULL) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    float& pk = SIMULST_EL(p_n, k);
                    if (is_in(k)) bad = bad || !(pk >= -1e-10f) || !(pk <= 1.0f);
                    if (!is_live(k)) pk = 0.f;
                    if (SOFT) {
                        if (!is_live(k)) SIMULST_EL(E_n, k) = is_in(k) ? fill : -INFINITY;
                    }
                }
            }
            xinc = local_cumprod<VPT>(p_n, eps, cpre);
            if (SOFT) {
                float2 unused[H], ex_n[H];
                einc = local_exp_prefix<VPT, false>(E_n, m_cur, eps_x, unused, ex_n, Dl);
                nan_out = nan_out || (einc != einc);
#pragma unroll
                for (int q = 0; q < VPT / 4; ++q)
                    stash[(sb_w * (VPT / 4) + q) * THREADS + tid] =
                        make_float4(ex_n[2 * q].x, ex_n[2 * q].y, ex_n[2 * q + 1].x, ex_n[2 * q + 1].y);
            }
        }
        // ---- RECU(it): local u-prefix
        float2 sl[H];
        float uinc = 0.f;
        if (doU) {
            uinc = local_u_prefix<VPT>(a_prev, rc, sl);
            nan_out = nan_out || (uinc != uinc);     // NaN anywhere in u poisons the thread total
        }
        // ---- warp level: x / e / u prefix scans, r suffix scan (RECR(it-1); its thread totals
        //      come from the previous POST), row max, row sum
        float rinc = rt_prev;
        float xexc, eexc, uexc, rexc;
        if (SOFT) {
            wscan_xeur(xinc, einc, uinc, rinc);
            wneigh_xeur(xinc, einc, uinc, rinc, xexc, eexc, uexc, rexc);
            const float wm = wmax_redux(em);
            if (lane == 0) xw[0 * kXStride + warp] = wm;
            if (lane == 31) xw[2 * kXStride + warp] = einc;
            if (lane == 0) xw[4 * kXStride + warp] = rinc;
        } else {
            wscan_xu(xinc, uinc);
            xexc = wprev(xinc, 1.f);
            uexc = wprev(uinc, 0.f);
            eexc = rexc = 0.f;
        }
        if (lane == 31) xw[1 * kXStride + warp] = xinc;
        if (lane == 31) xw[3 * kXStride + warp] = uinc;
        if (mp) {
            const float ws = warp_sum(rsum_prev);
            if (lane == 0) xw[5 * kXStride + warp] = ws;
        }
        if (want_d) {
            const float wd = warp_sum(wsum_prev);
            if (lane == 0) xw[6 * kXStride + warp] = wd;
        }

        __syncthreads();                    // ================================ the barrier
        // every thread has read ring slot `slotI`: refill it with the row NS steps ahead
        if (doI && it + 1 + NS < T_len) issue(it + 1 + NS, slotI);
        if (++slotI == NS) { slotI = 0; parI ^= 1u; }

        // ================================================================ POST
        if (SOFT && doM) m_cur = xw_max<NW>(xw + 0 * kXStride, lane);
        // ---- RECR(it-1): beta, mass-preserved column
        if (doR) {
            const int i = it - 1;
            float resid = 0.f, row_total = 0.f;
            if (mp) {
                row_total = xw_sum<NW>(xw + 5 * kXStride, lane);
                resid = 1.0f - fminf(fmaxf(row_total, 0.0f), 1.0f);
            }
            if (SOFT) {
                // R_k = sum_{q>=k} r_q  (+ resid / D_last for every k <= last; positions beyond
                // `last` are padded or outside the row, their beta is zero anyway)
                float rbase = xw_suffix_add<NW>(xw + 4 * kXStride, warp, lane) + rexc;
                if (mp) rbase += resid * bcast[i & 3];
                const float2 rb = f2(rbase);
                int sb_r = sb_w + 1;
                if (sb_r == kExStash) sb_r = 0;
                float2 b[H];
#pragma unroll
                for (int q = 0; q < VPT / 4; ++q) {
                    const float4 e4 = stash[(sb_r * (VPT / 4) + q) * THREADS + tid];
                    b[2 * q] = min2(mul2(f2(e4.x, e4.y), add2(rb, Rl[2 * q])), 1.0f);
                    b[2 * q + 1] = min2(mul2(f2(e4.z, e4.w), add2(rb, Rl[2 * q + 1])), 1.0f);
                }
                if constexpr (!FULL) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (!is_live(k)) SIMULST_EL(b, k) = 0.f;
                }
                if constexpr (MASKED) {
                    if (inside) st_row2_f32<VPT, FULL>(pb_run, 0, S, vec_out, b);
                    if (nl > 0 && nl < VPT) {
                        // the thread the row ends in: its columns beyond the row hold eps * R, not zero (threads
                        // wholly beyond the row compute exact zeros: their eps is zero) -- overwrite them
                        for (int k = nl; k < VPT; ++k) pb_run[k] = 0.f;
                    }
                } else {
                    if (inside) st_row2_f32<VPT, FULL>(g_beta + (size_t)i * ld_b, j0, S, vec_out, b);
                }
            }
            if (want_d) {
                // expected delay (mma_criterion.py:146-157): the weighted row sum left out the
                // column mass preservation rewrites (or holds its raw value when the residual is
                // ADDED), so the residual enters with that column's weight
                const float wtot = xw_sum<NW>(xw + 6 * kXStride, lane);
                if (tid == 0) prm.delays[(size_t)n * T_len + i] = mp ? wtot + (float)(last + 1) * resid : wtot;
            }
            if (own_last) {
                // the row itself was stored one iteration ago; patch the one column
                g_alpha[(size_t)i * ld_a + last] = mp_add ? (a_last_raw + resid) : resid;
                if (prm.side != nullptr)
                    *reinterpret_cast<float2*>(prm.side + ((size_t)n * T_len + i) * 2) = make_float2(a_last_raw, row_total);
            }
        }
        // ---- RECU(it): alpha_it, thread-local r-suffix / row sum of step it
        if (doU) {
            const float ubase = xw_prefix_add<NW>(xw + 3 * kXStride, warp, lane) + uexc;
            float2 sfull[H], z[H];
            finish_u_prefix<VPT>(ubase, sl, P, sfull, z);
#pragma unroll
            for (int q = 0; q < H; ++q) a_prev[q] = min2(z[q], 1.0f);      // z >= 0: P >= 0, s >= 0
            if constexpr (MASKED) {
                if (inside) st_row2_f32<VPT, FULL>(pa_run, 0, S, vec_out, a_prev);
            } else {
                if (inside) st_row2_f32<VPT, FULL>(g_alpha + (size_t)it * ld_a, j0, S, vec_out, a_prev);
            }
            if (mp || SOFT || want_d) {
                // alpha entering the row sum / the soft-attention numerator: the mass-preservation
                // column is left out when it is REPLACED (its residual is added analytically)
                float2 a_s[H];
#pragma unroll
                for (int q = 0; q < H; ++q) a_s[q] = a_prev[q];
                if constexpr (MASKED) {
                    if (own_last) a_last_raw = pick_el<VPT>(a_s, k_last);
                    if (!mp_add) {              // CTA-uniform: the no-mask rule on a SHIFT row (column REPLACED)
#pragma unroll
                        for (int k = 0; k < VPT; ++k)
                            if (own_last && k == k_last) SIMULST_EL(a_s, k) = 0.f;
                    }
                } else if (own_last) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (at_last(k)) {
                            a_last_raw = SIMULST_EL(a_s, k);
                            if (!mp_add) SIMULST_EL(a_s, k) = 0.f;
                        }
                }
                if (SOFT) rt_prev = local_r_suffix<VPT>(a_s, rD, Rl);
                if (mp) {
                    float2 acc = a_s[0];
#pragma unroll
                    for (int q = 1; q < H; ++q) acc = add2(acc, a_s[q]);
                    rsum_prev = acc.x + acc.y;
                }
                if (want_d) {
                    const float2 fj = f2((float)j0);
                    float2 acc = mul2(a_s[0], add2(fj, f2(1.0f, 2.0f)));
#pragma unroll
                    for (int q = 1; q < H; ++q) acc = fma2(a_s[q], add2(fj, f2((float)(2 * q + 1), (float)(2 * q + 2))), acc);
                    wsum_prev = acc.x + acc.y;
                }
            }
        }
        // ---- INV(it+1): cp, 1/c, P, 1/D
        if (doI) {
            const float xoff = xw_prefix_mul<NW>(xw + 1 * kXStride, warp, lane);
            const float cbase = (one_eps * xoff) * xexc;
            float2 cp[H];
            finish_cumprod<VPT>(cbase, cpre, p_n, eps, cp, rc, P);
            if (SOFT) {
                const float ebase = xw_prefix_add<NW>(xw + 2 * kXStride, warp, lane) + eexc;
                finish_exp_prefix<VPT>(ebase, eps, Dl, rD);
                if constexpr (MASKED) {
                    if (own_last) bcast[(it + 1) & 3] = pick_el<VPT>(rD, k_last);
                } else if (own_last) {
#pragma unroll
                    for (int k = 0; k < VPT; ++k)
                        if (at_last(k)) bcast[(it + 1) & 3] = SIMULST_EL(rD, k);
                }
            }
        }
        if (++sb_w == kExStash) sb_w = 0;
        if constexpr (MASKED) {
            pa_run += ld_a;
            if (SOFT) pb_run += ld_b;
        }
    };

    using Steady = std::integral_constant<bool, true>;
    using Edge = std::integral_constant<bool, false>;
    // iterations -2 .. T_len; all four stages are live for 1 <= it <= T_len - 3
    int it = -2;
    for (; it < 1 && it <= T_len; ++it) body(Edge{}, it);
    for (; it <= T_len - 3; ++it) body(Steady{}, it);
    for (; it <= T_len; ++it) body(Edge{}, it);

    // ---- data-error reporting (prob_check / safe_cumprod semantics), slow path only on error
    if (prm.status != nullptr) {
        if (nan_out) atomicOr(prm.status, SIMULST_ST_NAN);
        if (FULL) bad = SHIFT ? (umax32 > 0x3f800000u) : umax_trips<T>(umax);
        if (bad) {
            unsigned bits = 0u;
            for (int i = 0; i < T_len; ++i)
                for (int k = 0; k < VPT; ++k)
                    if (j0 + k < S) {
                        const float v = to_f32<T>(gp[(size_t)i * ld_p + j0 + k]);
                        bits |= prob_bits(v);
                        if ((1.0f - v) + eps < 0.f) bits |= SIMULST_ST_NEGPROD;
                    }
            if (bits) atomicOr(prm.status, bits);
        }
    }
}

// ------------------------------------------------------------------ host-side launcher
// Returns 1 when the row does not fit the pipelined kernel (caller falls back to the generic one).
template <int THREADS, int VPT, typename T, bool SOFT, bool FULL, bool DELAYS, bool RAGGED = false, bool MASKED = false,
          bool SHIFT = false>
int launch_mma_fwd_pipe_impl(const MmaParams& prm, cudaStream_t stream) {
    using PS = PipeStatic<THREADS, VPT, T, SOFT>;
    if (!PS::kFits) return 1;
    if (MASKED && PS::kNS < 2) return 1;        // the fixer looks two rows ahead
    auto kern = mma_fwd_pipe_kernel<THREADS, VPT, T, SOFT, FULL, DELAYS, RAGGED, MASKED, SHIFT>;
    static size_t attr_set[64] = {};    // per device: largest dynamic smem size opted into
    int dev = 0;
    cudaGetDevice(&dev);
    if (PS::kTotal > attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS::kTotal) != cudaSuccess) {
            cudaGetLastError();
            return SIMULST_E_LAUNCH;
        }
        attr_set[dev & 63] = PS::kTotal;
    }
    kern<<<prm.N, THREADS, PS::kTotal, stream>>>(prm);
    return check_launch();
}

// SOFT here means infinite lookback; requires prm.tma (16-byte aligned rows).
template <int THREADS, int VPT, typename T, bool SOFT>
int launch_mma_fwd_pipe(const MmaParams& prm, cudaStream_t stream) {
    // rows that are not 16-byte multiples / do not divide among the threads: shifted staging, live length
    // min(S, mask length); a mask needs the right-padding promise like the aligned masked kernel
    if (prm.shift && prm.S + 16 <= THREADS * VPT && !(prm.flags & SIMULST_MMA_LEFT_PADDING) &&
        (prm.mask == nullptr || (prm.flags & SIMULST_MMA_RIGHT_PADDING)))
        return launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true, true, true, true, true>(prm, stream);
    if (!prm.tma || prm.pitched) return 1;
    // right-padded rows (caller's promise): dense path with a per-row live length
    if (prm.mask != nullptr && (prm.flags & SIMULST_MMA_RIGHT_PADDING) && !(prm.flags & SIMULST_MMA_LEFT_PADDING) &&
        prm.vec_out && prm.S % VPT == 0 && prm.S <= THREADS * VPT)
        return launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true, true, true, true>(prm, stream);
    const bool dense = prm.mask == nullptr && prm.vec_out;
    const bool full = dense && prm.S == THREADS * VPT;
    // ragged dense rows: every thread wholly inside or wholly outside the row
    if (dense && !full && prm.S % VPT == 0 && prm.S < THREADS * VPT)
        return launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true, true, true>(prm, stream);
    if (!full) return launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, false, true>(prm, stream);
    return prm.delays != nullptr ? launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true, true>(prm, stream)
                                 : launch_mma_fwd_pipe_impl<THREADS, VPT, T, SOFT, true, false>(prm, stream);
}

}  // namespace simulst
