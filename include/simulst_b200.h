/*
 * simulst_b200 -- C ABI of the B200 (sm_100a) streaming-alignment hot path.
 *
 * Drop-in boundary for the tensor math of George0828Zhang/simulst (reference paths are
 * relative to its repository root).  The reference reaches this math through plain Python
 * functions; a maintainer binds the entry points below with ctypes (see INTEGRATION.md and
 * simulst_b200/_lib.py).  Everything is `extern "C"`, plain pointers and sizes.
 *
 * Conventions
 *  - All pointers are DEVICE pointers on the current CUDA device unless noted; tensors are
 *    dense row-major ("contiguous").  The caller allocates every input, output and
 *    workspace; the library never allocates or frees device memory and keeps no pointer
 *    after a call returns.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default
 *    stream).  No hidden synchronisation, no host reads: every call is CUDA-graph capturable.
 *  - Return value: 0 on success, a negative SIMULST_E_* code on argument / launch errors
 *    (checked on the host, synchronously).  No C++ exception crosses the ABI.
 *  - Data errors (NaN or out-of-range probabilities: the reference's `prob_check` /
 *    `safe_cumprod` assertions, codebase/utils/functions.py:9-17,57-61) are OR-ed by the
 *    kernels into the caller-provided device word `status` (may be NULL = don't record).
 *    The host wrapper decides when to look at it (strict: right away; default: lazily).
 *  - dtype enums select the element type of activation tensors; accumulation is always fp32.
 */
#ifndef SIMULST_B200_H_
#define SIMULST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMULST_VERSION 100 /* 0.1.0 */

/* element types */
#define SIMULST_F32 0
#define SIMULST_BF16 1
#define SIMULST_F16 2

/* error codes */
#define SIMULST_OK 0
#define SIMULST_E_ARG (-1)        /* null pointer / bad enum / bad flag combination */
#define SIMULST_E_SHAPE (-2)      /* dimension <= 0 or beyond the supported maximum */
#define SIMULST_E_ARCH (-3)       /* device is not sm_100 */
#define SIMULST_E_LAUNCH (-4)     /* cudaGetLastError() after launch */
#define SIMULST_E_ALIGN (-5)      /* pointer not aligned for its element type */

/* bits OR-ed into the device status word */
#define SIMULST_ST_NAN 1u         /* "Nan in a probability tensor."                    */
#define SIMULST_ST_RANGE 2u       /* "Incorrect values in a probability tensor"        */
#define SIMULST_ST_NEGPROD 4u     /* safe_cumprod: input + eps < 0                     */
#define SIMULST_ST_NOT_RIGHT_PADDED 8u /* SIMULST_MMA_RIGHT_PADDING was promised but a row's mask is not of
                                          the form (j >= len): results of that call are invalid */

/* flags of the MMA entry points */
#define SIMULST_MMA_MASS_PRESERVATION 1u /* apply mass_preservation to the alpha output  */
#define SIMULST_MMA_SOFT 2u              /* also produce beta (expected soft attention)  */
#define SIMULST_MMA_ENERGY_F16_FILL 4u   /* masked energy fill is -1e4 (fp16 energies)
                                            instead of -1e8 (monotonic_attention.py:106) */
#define SIMULST_MMA_LEFT_PADDING 8u      /* mass_preservation(left_padding=True)         */

#define SIMULST_MMA_RIGHT_PADDING 16u    /* caller's promise: padding_mask[n,j] = (j >= len_n), i.e.
                                            right padding (what mass_preservation(left_padding=False)
                                            and MMACriterion assume); lets masked rows take the dense
                                            backward kernel instead of the arbitrary-mask one */
#define SIMULST_MMA_MAX_SRC 16384        /* longest source row a single CTA keeps on chip */
#define SIMULST_SOFT_ATTENTION_MAX_SRC 9600 /* stand-alone simulst_soft_attention_*: 6 fp32 rows on chip */

int simulst_version(void);
const char* simulst_error_string(int code);

/* Number of kernels this library has launched since it was loaded / since the last reset
 * (host-side counter; used by bench.py for its `gpu_launches` claim). */
long long simulst_launch_count(void);
void simulst_reset_launch_count(void);

/* Tuning override for the MMA training kernels: threads per CTA and elements per thread
 * (threads * vpt >= S required).  0,0 restores automatic selection.  Returns 0 or E_ARG. */
int simulst_mma_set_config(int threads, int vpt);
/* 1 = stage rows with TMA bulk copies when alignment allows (default), 0 = cooperative loads */
int simulst_mma_set_tma(int enable);
/* Kernel choice for hard / infinite-lookback attention when rows can be TMA-staged.  Bit 0:
 * software-pipelined forward kernel; bit 2: dense fast-path backward kernel (rows that are
 * unmasked or right-padded and split evenly among the CTA's threads); cleared bits select the
 * generic (one scan per barrier) kernels.  Bit 1 is accepted and ignored (it selected a
 * pipelined backward that was slower than both alternatives and has been removed).  Default 5:
 * on B200 the pipelined forward is 1.3x and the fast backward 1.4x faster than the generic
 * kernels.  Returns 0 or E_ARG (mode outside 0..7). */
int simulst_mma_set_pipeline(int mode);
/* 1 (default): a call with a padding mask (and no LEFT/RIGHT_PADDING flag) runs as two passes
 * over disjoint row sets -- rows whose mask is a right-padding mask (j >= len) through the dense
 * kernels, every other row through the arbitrary-mask kernels; each CTA classifies its own row,
 * so there is no host read.  0: one pass through the arbitrary-mask kernels. */
int simulst_mma_set_mask_split(int enable);
/* Forward of long rows in small batches on thread-block clusters (csrc/mma_fwd_cluster.cuh): 2 / 4 / 8 CTAs
 * share a row of 1025..8192 frames (no padding mask, dense 16-byte rows) and exchange their scan totals
 * through distributed shared memory.  1 (default): when the call has at most 74 rows (half the SMs), i.e.
 * when one CTA per row would leave most of the GPU dark, and more than 2560 frames per row (below that the
 * cluster-wide exchange costs more than the extra SMs bring); 0: never; 2: whenever the shape qualifies.
 * Results are bit-identical to the single-CTA kernel's.  Returns 0 or E_ARG. */
int simulst_mma_set_cluster(int mode);
/* Development knob like simulst_mma_set_config: force the cluster shape (CTAs per row: 2 / 4 / 8; threads per
 * CTA: 96 / 128; eight frames per thread) of the calls that take the cluster kernel; a shape that cannot hold the
 * row is ignored for that call.  (0, 0) = automatic: slices of at most 1024 frames.  Returns 0 or E_ARG. */
int simulst_mma_set_cluster_shape(int cl, int threads);
/* 1 = CIF forward/backward through the TMA-staged tile kernels when rows are 16-byte aligned
 * and C <= 512 (default), 0 = always the per-warp kernels (same results bit for bit) */
int simulst_cif_set_tile(int enable);
/* Tuning override for the CIF tile kernels: frames staged per forward chunk and frames per
 * backward tile (multiple of 8); 0 = automatic (48 KB / 32 KB of shared memory).  Returns 0 or E_ARG. */
int simulst_cif_set_tile_rows(int fwd_chunk_frames, int bwd_tile_frames);

/* ---------------------------------------------------------------------------------------
 * MMA training path, forward.   Replaces, fused in one launch,
 *   expected_alignment_from_p_choose   codebase/utils/monotonic_attention.py:12-76
 *   mass_preservation                  codebase/utils/monotonic_attention.py:155-197
 *   expected_soft_attention            codebase/utils/monotonic_attention.py:79-152
 * i.e. steps 2-3 of MonotonicAttention.monotonic_attention_process_train
 * (codebase/modules/monotonic_multihead_attention.py:318-347).
 *
 *   p_choose      [N,T,S] p_dtype   stepwise probabilities (N = batch * heads)
 *   soft_energy   [N,T,S] e_dtype   soft-attention energies; NULL unless SIMULST_MMA_SOFT
 *   padding_mask  [N,S]   uint8     non-zero = padded source position; NULL = no padding
 *   alpha         [N,T,S] fp32 out  expected alignment (mass-preserved if the flag is set)
 *   beta          [N,T,S] fp32 out  expected soft attention; NULL unless SIMULST_MMA_SOFT
 *   side          [N,T,2] fp32 out  saved for backward when MASS_PRESERVATION is set:
 *                                   {alpha value overwritten by the residual, row sum};
 *                                   may be NULL when no backward will follow
 *   chunk_size    0 = infinite lookback, c >= 1 = chunkwise window (moving_sum)
 */
int simulst_mma_train_fwd(const void* p_choose, int p_dtype,
                          const void* soft_energy, int e_dtype,
                          const uint8_t* padding_mask,
                          float* alpha, float* beta, float* side,
                          int N, int T, int S,
                          float eps, int chunk_size, unsigned flags,
                          unsigned* status, void* stream);

/* MMA training path, backward (scans recomputed from p_choose / soft_energy; the only saved
 * tensors are the forward outputs).  Autograd of the three functions above.
 *   alpha, side     forward outputs
 *   grad_alpha      [N,T,S] fp32  dL/d(alpha output) or NULL (= 0)
 *   grad_beta       [N,T,S] fp32  dL/d(beta) or NULL (= 0; must be NULL without SOFT)
 *   grad_p          [N,T,S] gp_dtype out
 *   grad_energy     [N,T,S] ge_dtype out (NULL without SOFT)
 */
int simulst_mma_train_bwd(const void* p_choose, int p_dtype,
                          const void* soft_energy, int e_dtype,
                          const uint8_t* padding_mask,
                          const float* alpha, const float* side,
                          const float* grad_alpha, const float* grad_beta,
                          void* grad_p, int gp_dtype, void* grad_energy, int ge_dtype,
                          int N, int T, int S,
                          float eps, int chunk_size, unsigned flags,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * MMA training path with the expected-delay epilogue.  Same as
 * simulst_mma_train_fwd / _bwd plus
 *   expected_delays       [N,T] f32 out   sum_j (j+1) * alpha'_ij  of the OUTPUT (mass-preserved)
 *                                          row: the first step of the latency loss,
 *                                          codebase/criterion/mma_criterion.py:146-157
 *                                          (`steps * alpha_all` summed over the source axis),
 *                                          a by-product of the row scan instead of a second
 *                                          pass over alpha.  NULL: not produced.
 *   grad_expected_delays  [N,T] f32 in    dL/d expected_delays; the backward kernel adds
 *                                          gd_i * (j+1) to dL/d alpha'_ij on the fly, so a
 *                                          criterion that touches alpha only through the delays
 *                                          passes grad_alpha = NULL and the 4 B/element
 *                                          grad_alpha read disappears.  NULL: no such term.
 * Everything downstream of the delays (DifferentiableAverageLagging, head gathering, variance)
 * works on [N,T] tensors and stays in the caller. */
int simulst_mma_train_fwd_delays(const void* p_choose, int p_dtype,
                                 const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask,
                                 float* alpha, float* beta, float* side,
                                 float* expected_delays,
                                 int N, int T, int S,
                                 float eps, int chunk_size, unsigned flags,
                                 unsigned* status, void* stream);
int simulst_mma_train_bwd_delays(const void* p_choose, int p_dtype,
                                 const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask,
                                 const float* alpha, const float* side,
                                 const float* grad_alpha, const float* grad_beta,
                                 const float* grad_expected_delays,
                                 void* grad_p, int gp_dtype, void* grad_energy, int ge_dtype,
                                 int N, int T, int S,
                                 float eps, int chunk_size, unsigned flags,
                                 void* stream);

/* ---------------------------------------------------------------------------------------
 * MMA training path with ROW PITCHES ("element strides").  Same operators and
 * arguments as simulst_mma_train_{fwd,bwd}_delays; every [N,T,S] tensor additionally carries its
 * row pitch `ld_*` in ELEMENTS: row (n,t) starts at element (n*T + t) * ld (so the batch stride is
 * T * ld), ld >= S, 0 = dense (ld = S).  A tensor allocated as [N,T,ld] and used as x[..., :S] is
 * such a tensor.  The reference has no counterpart: torch ops take strided tensors implicitly
 * (codebase/utils/monotonic_attention.py:12-197 index [:, :, j]).
 *
 * Why it exists: rows are staged by 16-byte bulk copies, and a source length whose rows are not
 * 16-byte multiples (S = 1500 in bf16, S = 999) cannot give dense tensors 16-byte rows.  When the
 * OUTPUT tensors of a call (alpha, beta; grad_p, grad_energy) are 16-byte aligned with a pitch of
 * at least simulst_mma_out_pitch(S) elements, the dense kernels take the call whatever the
 * alignment of the INPUT rows: each input row is fetched as its 16-byte aligned superset and read
 * at its byte offset, the row's tail thread masks the columns beyond S, and the padding columns
 * [S, ld) of the outputs are written with zeros.  With dense outputs such shapes run on the generic
 * kernels (about 2.5x slower).  Padding masks, flags, status, errors: as in the dense entry points;
 * additionally SIMULST_E_SHAPE for a pitch below S.
 *
 * simulst_mma_out_pitch(S): smallest pitch (elements, a multiple of 8, >= S) that qualifies the
 * outputs of a call with source length S for the dense kernels; equals S whenever dense outputs
 * already qualify.  Negative SIMULST_E_* on a bad S. */
int simulst_mma_out_pitch(int S);
int simulst_mma_train_fwd_pitched(const void* p_choose, int p_dtype, long long ld_p,
                                  const void* soft_energy, int e_dtype, long long ld_e,
                                  const uint8_t* padding_mask,
                                  float* alpha, long long ld_alpha, float* beta, long long ld_beta,
                                  float* side, float* expected_delays,
                                  int N, int T, int S,
                                  float eps, int chunk_size, unsigned flags,
                                  unsigned* status, void* stream);
int simulst_mma_train_bwd_pitched(const void* p_choose, int p_dtype, long long ld_p,
                                  const void* soft_energy, int e_dtype, long long ld_e,
                                  const uint8_t* padding_mask,
                                  const float* alpha, long long ld_alpha, const float* side,
                                  const float* grad_alpha, long long ld_grad_alpha,
                                  const float* grad_beta, long long ld_grad_beta,
                                  const float* grad_expected_delays,
                                  void* grad_p, int gp_dtype, long long ld_grad_p,
                                  void* grad_energy, int ge_dtype, long long ld_grad_energy,
                                  int N, int T, int S,
                                  float eps, int chunk_size, unsigned flags,
                                  void* stream);

/* ---------------------------------------------------------------------------------------
 * Pooled p_choose producer: the MMA training path of the fixed pre-decision
 * wrappers.  Replaces FixedStrideMonotonicAttention.insert_zeros and the tail of its p_choose()
 * (codebase/modules/fixed_pre_decision.py:85-95 and :139-159) fused with the three functions of
 * simulst_mma_train_fwd: the caller hands over p_choose_pooled [N,T,Sp], Sp = ceil(S / ratio),
 * as produced by p_choose_from_qk on the pooled keys (:133-138), instead of the dense [N,T,S]
 * tensor in which only every ratio-th column (and column S-1) is non-zero:
 *     dense[n,t,j] = pooled[n,t,(j+1)/ratio - 1]   if (j+1) % ratio == 0
 *                  = pooled[n,t,Sp-1]              if j == S-1     (:156-159)
 *                  = 0                             otherwise
 * alpha is zero off that grid, so the T-sequential recurrence (and its backward) runs on
 * [N,T,Sp] only; the expected soft attention, whose rows are independent, and the dense alpha /
 * beta outputs are produced by streaming row kernels (csrc/mma_sparse.cu, four launches per
 * forward + backward).  That path serves: hard or infinite-lookback attention, S <= 4096 with
 * S*esize a multiple of 16 bytes, 16-byte aligned tensors, and either no padding mask or the
 * SIMULST_MMA_RIGHT_PADDING promise (forward() asserts right padding,
 * monotonic_multihead_attention.py:378-381; the kernels verify the promise per row) --
 * simulst_mma_pooled_is_fused() answers for a shape.  Every other shape is served by expanding
 * into `p_dense` and running the dense kernels.
 *
 *   p_pooled       [N,T,Sp] p_dtype
 *   p_dense        [N,T,S]  p_dtype out   the zero-upsampled p_choose the reference's
 *                                         process_train returns; optional (NULL) when the call is
 *                                         fused, required otherwise (it is then also the workspace)
 *   workspace      simulst_mma_pooled_workspace_bytes(N,T,S,ratio) bytes, 256-byte aligned,
 *                  written by the forward call and read by the backward call of the same step
 *                  (alpha on the grid and per-row geometry; scratch for the gradient on the grid).
 *                  NULL: the call takes the expand + dense path.
 *   grad_p_pooled  [N,T,Sp] gp_dtype out  gradient w.r.t. p_pooled (each pooled element owns one
 *                                         dense column)
 *   grad_p_dense   [N,T,S]  workspace     required (with p_dense) when the call is not fused
 * All other arguments as in simulst_mma_train_{fwd,bwd}_delays (`alpha` of the backward call is
 * not read on the fused path). */
int simulst_mma_pooled_is_fused(int p_dtype, int S, int ratio, int chunk_size, unsigned flags, int has_mask);
long long simulst_mma_pooled_workspace_bytes(int N, int T, int S, int ratio);
/* 1 (default): pooled calls use the pooled-grid kernels when the shape qualifies; 0: always expand
 * + dense kernels (development knob, same results within the parity tolerance) */
int simulst_mma_set_pooled_grid(int enable);
int simulst_mma_train_fwd_pooled(const void* p_pooled, int p_dtype, int ratio,
                                 const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, void* p_dense,
                                 float* alpha, float* beta, float* side, float* expected_delays,
                                 void* workspace,
                                 int N, int T, int S,
                                 float eps, int chunk_size, unsigned flags,
                                 unsigned* status, void* stream);
int simulst_mma_train_bwd_pooled(const void* p_pooled, int p_dtype, int ratio,
                                 const void* soft_energy, int e_dtype,
                                 const uint8_t* padding_mask, const void* p_dense,
                                 const float* alpha, const float* side,
                                 const float* grad_alpha, const float* grad_beta,
                                 const float* grad_expected_delays,
                                 void* grad_p_pooled, int gp_dtype, void* grad_p_dense,
                                 void* grad_energy, int ge_dtype,
                                 void* workspace,
                                 int N, int T, int S,
                                 float eps, int chunk_size, unsigned flags,
                                 void* stream);

/* ---------------------------------------------------------------------------------------
 * Gradient all-reduce of the training step over NVSwitch multicast (the path shards
 * by utterance with no data-path collective; the step it sits in sums parameter gradients once).
 * In-place SUM over `world` ranks of a symmetric fp32 buffer that has a multicast mapping: rank r
 * pulls the switch-reduced sum of its 1/world slice (multimem.ld_reduce) and pushes it to every
 * rank (multimem.st), from `ctas` CTAs of 512 threads -- a few CTAs instead of the SM footprint of
 * a ring all-reduce that overlaps compute.  The caller owns the symmetric allocation and the
 * cross-rank barriers before and after (torch.distributed._symmetric_memory: empty / rendezvous /
 * handle.multicast_ptr / handle.barrier); there is no reference counterpart (fairseq's DDP).
 *   multicast_ptr  multicast address of the buffer (16-byte aligned), numel a multiple of 4 */
int simulst_multimem_allreduce_f32(void* multicast_ptr, long long numel, int rank, int world,
                                   int ctas, void* stream);

/* ---------------------------------------------------------------------------------------
 * Latency loss next to the expected-delay epilogue.
 * DifferentiableAverageLagging as called by MMACriterion.compute_latency_loss
 * (codebase/criterion/mma_criterion.py:172-177) and CIFCriterion.compute_latency_loss
 * (codebase/criterion/cif_criterion.py:211-216); the function itself is SimulEval's
 * (simuleval.metrics.latency, not vendored by the reference):
 *     g'(0) = g(0), g'(i) = max(g'(i-1) + 1/gamma, g(i)), DAL = sum_i (g'(i) - i/gamma) / |Y|,
 *     gamma = ref_len / src_len (ref_lens given) or |Y| / src_len.
 *   delays               [N,T] fp32   expected delays (simulst_mma_train_fwd_delays / CIF delays)
 *   src_lens, ref_lens   [N] int64    ref_lens may be NULL
 *   target_padding_mask  [N,T] uint8  non-zero = padded target position; NULL = none
 *   dal                  [N] fp32 out
 * One warp per row, T sequential steps instead of ~3T launches.  T <= SIMULST_DAL_MAX_TGT. */
#define SIMULST_DAL_MAX_TGT 1024
int simulst_dal_fwd(const float* delays, const int64_t* src_lens, const int64_t* ref_lens,
                    const uint8_t* target_padding_mask, float* dal, int N, int T, void* stream);
/* grad_dal [N] fp32 -> grad_delays [N,T] fp32 (recomputes the recurrence). */
int simulst_dal_bwd(const float* delays, const int64_t* src_lens, const int64_t* ref_lens,
                    const uint8_t* target_padding_mask, const float* grad_dal,
                    float* grad_delays, int N, int T, void* stream);

/* ---------------------------------------------------------------------------------------
 * SSNT lattice loss.  Replaces ssnt_loss / ssnt_loss_mem
 * (codebase/criterion/ssnt_loss/ssnt_loss.py:45-151 / :154-271): the MMA recurrence in log space
 * with the word-prediction log-probability folded in,
 *     log_alpha[i+1] = clamp(trans[i] + log_p[i] + lcp[i]
 *                            + logcumsumexp(log1p(lambda) + log_alpha[i] - lcp[i]), neg_inf, 0).
 * One CTA per sample, lattice row on chip, scans over the source axis; backward recomputes them.
 *
 *   log_probs        [rows, S, V] lp_dtype  word log-probs (log_softmax output); rows = N*T in the
 *                                           padded layout, T_flat (targets concatenated) in the flat one
 *   targets          [rows] int64
 *   emit             [rows, S] e_dtype      emission logits (emit_is_logits != 0) or probabilities
 *   source_lengths, target_lengths [N] int64
 *   row_offsets      [N] int64   flat layout: first row of sample n (exclusive cumsum of target_lengths);
 *                                NULL selects the padded layout (sample n owns rows n*T .. n*T+T-1)
 *   lattice_offsets  [N] int64   flat layout: row of sample n's alpha_0 in `lattice`
 *                                (exclusive cumsum of target_lengths + 1); NULL in the padded layout
 *   lattice          fp32 out    padded: [N, T, S] = log_alpha[:, 1:]; flat: [T_flat + N, S] incl. alpha_0 rows
 *   log_p_choose     [rows, S] fp32 out  log p with source padding filled with neg_inf (may be NULL)
 *   loss             [N] fp32 out        -log_alpha[n, target_len, source_len - 1] (reduction: caller)
 * S <= SIMULST_SSNT_MAX_SRC.  NaN in the lattice sets SIMULST_ST_NAN in `status`. */
#define SIMULST_SSNT_MAX_SRC 4096
int simulst_ssnt_fwd(const void* log_probs, int lp_dtype, const int64_t* targets,
                     const void* emit, int e_dtype, int emit_is_logits,
                     const int64_t* source_lengths, const int64_t* target_lengths,
                     const int64_t* row_offsets, const int64_t* lattice_offsets,
                     float* lattice, float* log_p_choose, float* loss,
                     int N, int T, int S, int V, float neg_inf, float fastemit_lambda,
                     unsigned* status, void* stream);
/* Backward.  grad_loss [N]; grad_lattice (layout of lattice) and grad_log_p_choose [rows, S] may be
 * NULL.  grad_emit [rows, S] e_dtype out.  grad_log_probs [rows, S, V] lp_dtype must be ZERO-FILLED
 * by the caller (NULL: not wanted): the kernel writes the one gathered column per (row, frame). */
int simulst_ssnt_bwd(const void* log_probs, int lp_dtype, const int64_t* targets,
                     const void* emit, int e_dtype, int emit_is_logits,
                     const int64_t* source_lengths, const int64_t* target_lengths,
                     const int64_t* row_offsets, const int64_t* lattice_offsets,
                     const float* lattice, const float* grad_loss, const float* grad_lattice,
                     const float* grad_log_p_choose, void* grad_emit, void* grad_log_probs,
                     int N, int T, int S, int V, float neg_inf, float fastemit_lambda, void* stream);
/* prob_check(log_probs, neg_inf, logp=True) (ssnt_loss.py:29-42, called at :80 and :200): NaN ->
 * SIMULST_ST_NAN, value > 0 or < neg_inf -> SIMULST_ST_RANGE, OR-ed into `status`; one streaming
 * pass over the tensor (16-byte aligned pointer). */
int simulst_logprob_check(const void* log_probs, int dtype, long long numel, float neg_inf,
                          unsigned* status, void* stream);

/* ---------------------------------------------------------------------------------------
 * CTC best alignment.  Replaces the reference's only native kernel and the
 * Python around it, in one launch:
 *   ctc_alignment_log_alpha_gpu_kernel   codebase/criterion/best_alignment/best_alignment.cu:58-202
 *   best_alignment(...)                  codebase/criterion/best_alignment/__init__.py:25-111
 *                                        (final-state choice, S-iteration back-trace, labels)
 *   log_probs       [S, N, V] dtype   log emission probabilities (after log_softmax), time-major
 *   targets         [N, target_stride] int64, target_stride >= Tmax
 *   input_lengths, target_lengths [N] int64  (device pointers: the reference copies them to the
 *                                     host and back, best_alignment.cpp:16-23)
 *   workspace       simulst_ctc_workspace_bytes(N, S, Tmax) bytes: one BYTE per (sample, frame,
 *                   state) holding the arg-max jump 0/1/2 (the reference stores int64 indices and a
 *                   fp32 log_alpha of the same shape)
 *   nll             [N] fp32 out, may be NULL   -log(sum of the two final states), .cu:187-201
 *   states          [N, S] int64 out   state sequence in [0, 2T+1); 0 for frames >= input_length
 *   labels          [N, S] int64 out, may be NULL   states translated to labels (as_labels=True)
 * 2*Tmax+1 <= 4096.  No host read: CUDA-graph capturable. */
long long simulst_ctc_workspace_bytes(int N, int S, int Tmax);
int simulst_ctc_best_alignment(const void* log_probs, int dtype,
                               const int64_t* targets, int target_stride,
                               const int64_t* input_lengths, const int64_t* target_lengths,
                               int blank, uint8_t* workspace, float* nll,
                               int64_t* states, int64_t* labels,
                               int N, int S, int V, int Tmax, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stand-alone pieces (same math, rows independent): used when the reference functions are
 * called one by one rather than through monotonic_attention_process_train.
 */

/* expected_soft_attention(alpha, soft_energy, padding_mask, chunk_size, eps)
 * (monotonic_attention.py:79-152).  a_dtype is the dtype of alpha AND of the beta output. */
int simulst_soft_attention_fwd(const void* alpha, int a_dtype,
                               const void* soft_energy, int e_dtype,
                               const uint8_t* padding_mask, void* beta,
                               int N, int T, int S, float eps, int chunk_size, unsigned flags,
                               unsigned* status, void* stream);
int simulst_soft_attention_bwd(const void* alpha, int a_dtype,
                               const void* soft_energy, int e_dtype,
                               const uint8_t* padding_mask, const void* grad_beta,
                               void* grad_alpha, void* grad_energy,
                               int N, int T, int S, float eps, int chunk_size, unsigned flags,
                               void* stream);

/* mass_preservation(alpha, padding_mask, left_padding) (monotonic_attention.py:155-197).
 * In place on `alpha` [N,T,S] fp32; writes side [N,T,2] like the fused forward. */
int simulst_mass_preservation_fwd(float* alpha, const uint8_t* padding_mask, float* side,
                                  int N, int T, int S, unsigned flags,
                                  unsigned* status, void* stream);
/* grad_in [N,T,S] fp32 (dL/d output) -> grad_out [N,T,S] fp32 (dL/d input); may alias. */
int simulst_mass_preservation_bwd(const float* grad_in, const uint8_t* padding_mask,
                                  const float* side, float* grad_out,
                                  int N, int T, int S, unsigned flags, void* stream);

/* moving_sum(x, start_idx, end_idx) (codebase/utils/functions.py:69-125):
 * out[n] = sum_{m = n-start+1}^{n+end-1} x[m] along the last axis, zero outside. rows = N*T. */
int simulst_moving_sum(const void* x, void* out, int dtype, long long rows, int S,
                       int start_idx, int end_idx, void* stream);

/* exclusive_cumprod(x, dim=last, eps) = exp(cumsum(log(cat[1, x] + eps)))[:-1]
 * (functions.py:20-45); inclusive != 0 gives safe_cumprod = exp(cumsum(log(x + eps)))
 * (functions.py:48-66). */
int simulst_exclusive_cumprod(const void* x, void* out, int dtype, long long rows, int S,
                              float eps, int inclusive, unsigned* status, void* stream);

/* Autograd of the two functions above (the reference composes log / cumsum / exp, which torch
 * differentiates): grad_x_k = sum_{j >= k + (inclusive ? 0 : 1)} grad_y_j * y_j / (x_k + eps),
 * y = the forward output.  The backward of moving_sum is moving_sum itself with start_idx and
 * end_idx exchanged. */
int simulst_cumprod_bwd(const void* x, const void* y, const void* grad_y, void* grad_x, int dtype,
                        long long rows, int S, float eps, int inclusive, void* stream);

/* learnable_p_choose (codebase/utils/p_choose_strategy.py:56-76):
 * out = sigmoid(energy + noise), noise may be NULL (eval).  Same dtype in and out; the
 * Gaussian noise is drawn by the caller (torch RNG) so that results are reproducible. */
int simulst_p_choose(const void* energy, const void* noise, void* out, int dtype,
                     long long numel, void* stream);

/* ---------------------------------------------------------------------------------------
 * MMA incremental decoding step.  Replaces the body of
 * MonotonicAttention.monotonic_attention_process_infer
 * (codebase/modules/monotonic_multihead_attention.py:171-299) after the energy bmm's.
 *
 *   p_choose     [R,S] p_dtype      sigmoid(monotonic energy) for this step (R = bsz*heads)
 *   soft_energy  [R,S] e_dtype      NULL for hard-aligned attention
 *   src_lengths  [R] int32          valid source length per row; NULL = S for all rows
 *   head_step    [R] int64 in/out   the `head_step` cache (zeros before the first step)
 *   head_read    [R] uint8 out      the `head_read` cache (bool)
 *   alpha        [R,S] p_dtype out  one-hot alignment (zeros_like(p_choose), :261)
 *   beta         [R,S] e_dtype out  softmax over the look-back window; NULL for hard
 *   flags        SIMULST_MMA_MASS_PRESERVATION or 0
 */
int simulst_mma_step(const void* p_choose, int p_dtype,
                     const void* soft_energy, int e_dtype,
                     const int32_t* src_lengths,
                     int64_t* head_step, uint8_t* head_read,
                     void* alpha, void* beta,
                     int R, int S, unsigned flags, void* stream);

/* Sizes the caller needs for the CIF calls below (the library never allocates):
 *   simulst_cif_workspace_bytes(B, S)   bytes of the `workspace` argument of simulst_cif_bwd
 *   simulst_cif_seg_stride(training, t_cap, S, beta)   minimum row stride (in int32) of `seg_first`:
 *       training: t_cap + 2 with t_cap = max target length; inference: floor(S / beta) + 3 */
long long simulst_cif_workspace_bytes(int B, int S);
int simulst_cif_seg_stride(int training, int t_cap, int S, float beta);

/* ---------------------------------------------------------------------------------------
 * CIF (continuous integrate-and-fire).  Replaces cif_function
 * (codebase/models/torch_cif/cif.py:23-196) with a plan pass + a gather pass (+ two backward
 * passes); the host reads the output length exactly where the reference does
 * (`feat_lengths.max()`, cif.py:72/76 and :181).
 *
 * Pass 1, simulst_cif_plan: per batch row, mask + scale alpha (training mode), inclusive scan.
 *   alpha          [B,S] a_dtype   integration weights (after sigmoid)
 *   padding_mask   [B,S] uint8     or NULL
 *   desired_sum    [B] fp32        training: beta*target_length+eps evaluated by the host
 *                                  wrapper in the input dtype (cif.py:68); NULL = inference
 *   target_lengths [B] int64       training lengths; NULL = inference
 *   csum           [B,S] fp32 out  cumsum of the (scaled, masked) weights (cif.py:79)
 *   scale          [B] fp32 out    desired_sum / alpha_sum (1 in inference)
 *   alpha_sum      [B] fp32 out    sum of the masked, unscaled weights (cif.py:69,74)
 *   lengths        [B] int64 out   feat_lengths before tail handling
 *   t_max          [1] int32 i/o   inference: atomicMax of lengths (caller zeroes it); may be
 *                                  NULL in training
 *   seg_first      [B,seg_stride] int32 out  segment table: seg_first[b,t] = first frame whose
 *                                  firing index floor(csum/beta) reaches t (S if none), so that
 *                                  output slot t draws from frames seg_first[t]..seg_first[t+1].
 *                                  seg_stride >= T+2 (training) / floor(S/beta)+3 (inference)
 */
int simulst_cif_plan(const void* alpha, int a_dtype, const uint8_t* padding_mask,
                     const float* desired_sum, const int64_t* target_lengths,
                     float* csum, float* scale, float* alpha_sum, int64_t* lengths,
                     int* t_max, int32_t* seg_first, int seg_stride, int B, int S, float beta,
                     unsigned* status, void* stream);

/* Pass 2, simulst_cif_fwd: weighted segment sums, one warp per output slot (every slot reads
 * one contiguous source range; deterministic, no atomics), fused with the inference tail
 * handling (cif.py:156-188).
 *   input        [B,S,C] x_dtype
 *   T            max over rows of `lengths` (firing indices are clipped to it, cif.py:82)
 *   T_alloc      slots computed per row: T in training, T+1 in inference
 *   seg_first    [B,seg_stride] int32   the table written by simulst_cif_plan
 *   cif_out      [B,T_alloc,C] x_dtype out     delays [B,T_alloc] x_dtype out
 *   tail_weights [B] fp32 out      (inference)  lengths_out [B] int64 out (inference: lengths
 *                                  + 1 where the tail fires)   t_max2 [1] int32 i/o: atomicMax of
 *                                  lengths_out (caller zeroes it)
 */
int simulst_cif_fwd(const void* input, int x_dtype, const float* csum, const float* scale,
                    const void* alpha, int a_dtype, const uint8_t* padding_mask,
                    const int32_t* seg_first, int seg_stride,
                    void* cif_out, void* delays, float* tail_weights,
                    const int64_t* lengths, int64_t* lengths_out, int* t_max2,
                    int B, int S, int C, int T, int T_alloc,
                    float beta, float tail_thres, int training, void* stream);

/* CIF backward: dL/d input [B,S,C] and dL/d alpha [B,S] given dL/d cif_out [B,T_out,C],
 * dL/d delays [B,T_out] (may be NULL) and dL/d alpha_sum [B] (may be NULL).  T_out is the
 * number of slots the caller kept (T in training; max(lengths_after_tail) in inference).
 * workspace: 2*B*S floats. */
int simulst_cif_bwd(const void* input, int x_dtype, const float* csum, const float* scale,
                    const void* alpha, int a_dtype, const uint8_t* padding_mask,
                    const void* grad_out, const void* grad_delays,
                    const float* tail_weights, const int64_t* lengths_before_tail,
                    const int64_t* lengths_after_tail,
                    const float* alpha_sum, const float* grad_alpha_sum,
                    void* grad_input, void* grad_alpha, float* workspace,
                    int B, int S, int C, int T, int T_out,
                    float beta, float tail_thres, int training, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMULST_B200_H_ */
