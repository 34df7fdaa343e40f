"""torch.autograd wrappers over the C ABI (one Function per fused operator).

Each Function allocates outputs with torch, passes raw device pointers + the current CUDA
stream to libsimulst_b200.so and saves only what the recompute-based backward needs.
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def _mask_u8(mask: Optional[Tensor], n: int, s: int, device) -> Optional[Tensor]:
    if mask is None:
        return None
    if tuple(mask.shape) != (n, s):
        raise ValueError(f"padding_mask must be [{n}, {s}], got {tuple(mask.shape)}")
    m = mask.to(device=device)
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    return (m != 0).contiguous().view(torch.uint8)


def _rows(x: Tensor):
    """A [N,T,S] tensor as (tensor, row pitch in elements) for the pitched entry points: dense tensors and
    tensors whose rows are unit-stride with a uniform pitch (e.g. x_padded[..., :S]) pass as they are,
    anything else is made contiguous first."""
    n, t, s = x.shape
    if x.is_contiguous():
        return x, s
    if x.stride(2) == 1 and x.stride(1) >= s and x.stride(0) == t * x.stride(1):
        return x, x.stride(1)
    return x.contiguous(), s


def _empty_rows(n: int, t: int, s: int, ld: int, dtype, dev):
    """[N,T,S] output with row pitch ld: the [..., :S] view of a [N,T,ld] buffer when ld > S."""
    if ld == s:
        return torch.empty((n, t, s), dtype=dtype, device=dev)
    return torch.empty((n, t, ld), dtype=dtype, device=dev)[..., :s]


class MMATrainFunction(torch.autograd.Function):
    """(p_choose, soft_energy) -> (alpha, beta): steps 2-3 of
    MonotonicAttention.monotonic_attention_process_train
    (reference codebase/modules/monotonic_multihead_attention.py:318-347) in one launch;
    backward is one launch that recomputes every scan."""

    @staticmethod
    def forward(ctx, p_choose: Tensor, soft_energy: Optional[Tensor], padding_mask: Optional[Tensor],
                eps: float, mass_preservation: bool, chunk_size: Optional[int],
                left_padding: bool = False, with_delays: bool = False):
        lib = _lib.load()
        dev = _lib.require_cuda(p_choose, soft_energy, padding_mask)
        if p_choose.dim() != 3:
            raise ValueError("p_choose must be [bsz*heads, tgt_len, src_len]")
        n, t, s = p_choose.shape
        if s > _lib.MMA_MAX_SRC:
            raise ValueError(f"src_len {s} exceeds the on-chip row limit {_lib.MMA_MAX_SRC}")
        soft = soft_energy is not None
        p, e = p_choose, None
        flags = 0
        if soft:
            if tuple(soft_energy.shape) != (n, t, s):
                raise ValueError("soft_energy must have the shape of p_choose")
            if soft_energy.dtype == torch.float16:
                flags |= _lib.MMA_ENERGY_F16_FILL      # -1e4 fill, monotonic_attention.py:106
            e = soft_energy
            if e.dtype != p.dtype:                      # kernel wants one activation dtype
                common = torch.promote_types(e.dtype, p.dtype)
                e, p = e.to(common), p.to(common)
            flags |= _lib.MMA_SOFT
            e, ld_e = _rows(e)
        else:
            ld_e = 0
        p, ld_p = _rows(p)
        if mass_preservation:
            flags |= _lib.MMA_MASS_PRESERVATION
        if left_padding:
            flags |= _lib.MMA_LEFT_PADDING
        mask = _mask_u8(padding_mask, n, s, dev)
        if mask is not None and not left_padding and _lib.right_padding_assumed():
            flags |= _lib.MMA_RIGHT_PADDING
        # rows that are not 16-byte multiples: 16-byte pitched outputs keep them on the dense kernels
        ld_out = int(lib.simulst_mma_out_pitch(s)) if _lib.pitched_outputs() else s
        alpha = _empty_rows(n, t, s, ld_out, torch.float32, dev)
        beta = _empty_rows(n, t, s, ld_out, torch.float32, dev) if soft else None
        # a row whose mask leaves no live column has no mass-preservation column: the kernels
        # write nothing for it, so with a mask these two small buffers start as zeros
        small = torch.zeros if mask is not None else torch.empty
        side = small((n, t, 2), dtype=torch.float32, device=dev) if mass_preservation else None
        delays = small((n, t), dtype=torch.float32, device=dev) if with_delays else None
        status = _lib.status_word(dev)
        chunk = int(chunk_size) if chunk_size else 0
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_fwd_pitched(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), ld_p, _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, ld_e, _lib.ptr(mask),
                _lib.ptr(alpha), ld_out, _lib.ptr(beta), ld_out, _lib.ptr(side), _lib.ptr(delays),
                n, t, s, float(eps), chunk, flags, _lib.ptr(status), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_fwd_pitched")
        _lib.maybe_check(dev)
        ctx.save_for_backward(p, e, mask, alpha, side)
        ctx.cfg = (n, t, s, float(eps), chunk, flags, soft)
        ctx.pitches = (ld_p, ld_e, ld_out)
        ctx.in_dtypes = (p_choose.dtype, soft_energy.dtype if soft else None)
        ctx.mark_non_differentiable()
        ctx.set_materialize_grads(False)        # an unused output costs no zero-filled gradient
        return (alpha, beta if soft else alpha.new_empty(0),
                delays if with_delays else alpha.new_empty(0))

    @staticmethod
    def backward(ctx, g_alpha, g_beta, g_delays):
        lib = _lib.load()
        p, e, mask, alpha, side = ctx.saved_tensors
        n, t, s, eps, chunk, flags, soft = ctx.cfg
        dev = p.device
        ld_p, ld_e, ld_out = ctx.pitches
        ga, ld_ga = _rows(g_alpha.float()) if g_alpha is not None else (None, 0)
        gb, ld_gb = _rows(g_beta.float()) if (soft and g_beta is not None) else (None, 0)
        gd = g_delays.contiguous().float() if (g_delays is not None and g_delays.numel() == n * t) else None
        grad_p = _empty_rows(n, t, s, ld_out, p.dtype, dev)
        grad_e = _empty_rows(n, t, s, ld_out, e.dtype, dev) if soft else None
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_bwd_pitched(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), ld_p, _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, ld_e, _lib.ptr(mask),
                _lib.ptr(alpha), ld_out, _lib.ptr(side), _lib.ptr(ga), ld_ga, _lib.ptr(gb), ld_gb, _lib.ptr(gd),
                _lib.ptr(grad_p), _lib.dtype_enum(p.dtype), ld_out, _lib.ptr(grad_e),
                _lib.dtype_enum(e.dtype) if soft else 0, ld_out,
                n, t, s, eps, chunk, flags, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_bwd_pitched")
        p_dt, e_dt = ctx.in_dtypes
        if grad_p.dtype != p_dt:
            grad_p = grad_p.to(p_dt)
        if soft and grad_e.dtype != e_dt:
            grad_e = grad_e.to(e_dt)
        return grad_p, grad_e, None, None, None, None, None, None


def mma_train(p_choose: Tensor, soft_energy: Optional[Tensor] = None,
              padding_mask: Optional[Tensor] = None, eps: float = 1e-6,
              mass_preservation: bool = True, chunk_size: Optional[int] = None,
              left_padding: bool = False):
    """Fused expected alignment (+ mass preservation) (+ expected soft attention).
    Returns (alpha [N,T,S] fp32, beta [N,T,S] fp32); beta is alpha for hard attention."""
    alpha, beta, _ = MMATrainFunction.apply(p_choose, soft_energy, padding_mask, eps,
                                            mass_preservation, chunk_size, left_padding, False)
    if soft_energy is None:
        beta = alpha
    return alpha, beta


def mma_train_with_delays(p_choose: Tensor, soft_energy: Optional[Tensor] = None,
                          padding_mask: Optional[Tensor] = None, eps: float = 1e-6,
                          mass_preservation: bool = True, chunk_size: Optional[int] = None,
                          left_padding: bool = False):
    """mma_train plus the expected delays `sum_j (j+1) * alpha[n,i,j]` ([N,T] fp32) of the
    returned alpha: step 2 of MMACriterion.compute_latency_loss
    (reference codebase/criterion/mma_criterion.py:146-157) as a by-product of the alignment
    kernel.  Differentiable: a loss that reaches alpha only through the delays (and beta) makes
    the backward kernel skip the [N,T,S] grad_alpha read entirely.
    Returns (alpha, beta, expected_delays)."""
    alpha, beta, delays = MMATrainFunction.apply(p_choose, soft_energy, padding_mask, eps,
                                                 mass_preservation, chunk_size, left_padding, True)
    if soft_energy is None:
        beta = alpha
    return alpha, beta, delays


# ----------------------------------------------------------------------------- pooled p_choose (fixed pre-decision)
class MMATrainPooledFunction(torch.autograd.Function):
    """(p_choose_pooled [N,T,ceil(S/ratio)], soft_energy [N,T,S]) -> (alpha, beta, delays, p_choose):
    FixedStrideMonotonicAttention.insert_zeros + the tail of its p_choose()
    (reference codebase/modules/fixed_pre_decision.py:85-95, :139-159) fused with steps 2-3 of
    monotonic_attention_process_train.  The zero-upsampled row is formed in registers; the dense
    p_choose is written only when ``want_dense`` (the reference returns it in the attention dict,
    nothing differentiates through it: it comes back non-differentiable)."""

    @staticmethod
    def forward(ctx, p_pooled: Tensor, soft_energy: Optional[Tensor], padding_mask: Optional[Tensor],
                src_len: int, ratio: int, eps: float, mass_preservation: bool, chunk_size: Optional[int],
                with_delays: bool, want_dense: bool, right_padding: bool, want_alpha: bool = True):
        lib = _lib.load()
        dev = _lib.require_cuda(p_pooled, soft_energy, padding_mask)
        if p_pooled.dim() != 3:
            raise ValueError("p_choose_pooled must be [bsz*heads, tgt_len, ceil(src_len / ratio)]")
        n, t, sp = p_pooled.shape
        s, ratio = int(src_len), int(ratio)
        if ratio < 2 or sp != (s + ratio - 1) // ratio:
            raise ValueError(f"p_choose_pooled has {sp} columns, expected ceil({s} / {ratio})")
        if s > _lib.MMA_MAX_SRC:
            raise ValueError(f"src_len {s} exceeds the on-chip row limit {_lib.MMA_MAX_SRC}")
        soft = soft_energy is not None
        p = p_pooled.contiguous()
        e = None
        flags = 0
        if soft:
            if tuple(soft_energy.shape) != (n, t, s):
                raise ValueError("soft_energy must be [bsz*heads, tgt_len, src_len]")
            if soft_energy.dtype == torch.float16:
                flags |= _lib.MMA_ENERGY_F16_FILL
            e = soft_energy.contiguous()
            if e.dtype != p.dtype:
                common = torch.promote_types(e.dtype, p.dtype)
                e, p = e.to(common), p.to(common)
            flags |= _lib.MMA_SOFT
        if mass_preservation:
            flags |= _lib.MMA_MASS_PRESERVATION
        mask = _mask_u8(padding_mask, n, s, dev)
        if mask is not None and (right_padding or _lib.right_padding_assumed()):
            flags |= _lib.MMA_RIGHT_PADDING
        chunk = int(chunk_size) if chunk_size else 0
        fused = bool(lib.simulst_mma_pooled_is_fused(_lib.dtype_enum(p.dtype), s, ratio, chunk, flags,
                                                     1 if mask is not None else 0))
        need_dense = want_dense or not fused
        p_dense = torch.empty((n, t, s), dtype=p.dtype, device=dev) if need_dense else None
        ws = None
        if fused:       # alpha on the grid + row geometry, carried to the backward call
            ws = torch.empty(int(lib.simulst_mma_pooled_workspace_bytes(n, t, s, ratio)), dtype=torch.uint8,
                             device=dev)
        # the dense alpha can be skipped on the pooled-grid path (soft attention, no dense p_choose):
        # the caller then reaches alpha through beta and the expected delays only
        skip_alpha = (not want_alpha) and fused and soft and not need_dense
        alpha = None if skip_alpha else torch.empty((n, t, s), dtype=torch.float32, device=dev)
        beta = torch.empty((n, t, s), dtype=torch.float32, device=dev) if soft else None
        small = torch.zeros if mask is not None else torch.empty
        side = small((n, t, 2), dtype=torch.float32, device=dev) if mass_preservation else None
        delays = small((n, t), dtype=torch.float32, device=dev) if with_delays else None
        status = _lib.status_word(dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_fwd_pooled(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), ratio, _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, _lib.ptr(mask), _lib.ptr(p_dense),
                _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(side), _lib.ptr(delays), _lib.ptr(ws),
                n, t, s, float(eps), chunk, flags, _lib.ptr(status), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_fwd_pooled")
        _lib.maybe_check(dev)
        ctx.save_for_backward(p, e, mask, alpha, side, None if fused else p_dense, ws)
        ctx.cfg = (n, t, s, ratio, float(eps), chunk, flags, soft, fused)
        ctx.in_dtypes = (p_pooled.dtype, soft_energy.dtype if soft else None)
        ctx.set_materialize_grads(False)
        empty = p.new_empty(0, dtype=torch.float32)
        out_dense = p_dense if want_dense else empty.to(p_pooled.dtype)
        if out_dense.dtype != p_pooled.dtype:
            out_dense = out_dense.to(p_pooled.dtype)
        ctx.mark_non_differentiable(out_dense)
        return (alpha if alpha is not None else empty.clone(), beta if soft else empty.clone(),
                delays if with_delays else empty.clone(), out_dense)

    @staticmethod
    def backward(ctx, g_alpha, g_beta, g_delays, _g_dense):
        lib = _lib.load()
        p, e, mask, alpha, side, p_dense, ws = ctx.saved_tensors
        n, t, s, ratio, eps, chunk, flags, soft, fused = ctx.cfg
        dev = p.device
        ga = g_alpha.contiguous().float() if (g_alpha is not None and g_alpha.numel() == n * t * s) else None
        gb = g_beta.contiguous().float() if (soft and g_beta is not None) else None
        gd = g_delays.contiguous().float() if (g_delays is not None and g_delays.numel() == n * t) else None
        grad_p = torch.empty_like(p)
        grad_e = torch.empty_like(e) if soft else None
        grad_dense = None if fused else torch.empty((n, t, s), dtype=p.dtype, device=dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_bwd_pooled(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), ratio, _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, _lib.ptr(mask), _lib.ptr(p_dense),
                _lib.ptr(alpha), _lib.ptr(side), _lib.ptr(ga), _lib.ptr(gb), _lib.ptr(gd),
                _lib.ptr(grad_p), _lib.dtype_enum(p.dtype), _lib.ptr(grad_dense), _lib.ptr(grad_e),
                _lib.dtype_enum(e.dtype) if soft else 0, _lib.ptr(ws),
                n, t, s, eps, chunk, flags, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_bwd_pooled")
        p_dt, e_dt = ctx.in_dtypes
        if grad_p.dtype != p_dt:
            grad_p = grad_p.to(p_dt)
        if soft and grad_e.dtype != e_dt:
            grad_e = grad_e.to(e_dt)
        return (grad_p, grad_e) + (None,) * 10


def mma_train_pooled(p_choose_pooled: Tensor, src_len: int, ratio: int,
                     soft_energy: Optional[Tensor] = None, padding_mask: Optional[Tensor] = None,
                     eps: float = 1e-6, mass_preservation: bool = True, chunk_size: Optional[int] = None,
                     with_delays: bool = False, want_dense: bool = True, right_padding: bool = False,
                     want_alpha: bool = True):
    """Fixed pre-decision training path from the POOLED p_choose.
    Returns (p_choose [N,T,S] or None, alpha, beta, expected_delays or None); beta is alpha for hard
    attention.  ``right_padding=True`` is the caller's promise that ``padding_mask`` is a
    right-padding mask (verified on the device: a violation poisons the outputs with NaN and sets
    SIMULST_ST_NOT_RIGHT_PADDED); without it a masked call expands the row and runs the dense path.
    ``want_alpha=False`` (with ``want_dense=False``, soft attention, a shape the pooled-grid kernels
    take) skips the dense [N,T,S] alpha: alpha then reaches the caller through beta and the expected
    delays only, which is all MMACriterion uses it for (mma_criterion.py:146-157)."""
    alpha, beta, delays, dense = MMATrainPooledFunction.apply(
        p_choose_pooled, soft_energy, padding_mask, src_len, ratio, eps, mass_preservation, chunk_size,
        with_delays, want_dense, right_padding, want_alpha)
    if soft_energy is None:
        beta = alpha
    if alpha.numel() == 0 and p_choose_pooled.numel() != 0:
        alpha = None        # want_alpha=False on the pooled-grid path: the dense alpha was never written
    return (dense if want_dense else None), alpha, beta, (delays if with_delays else None)


# ----------------------------------------------------------------------------- stand-alone MMA pieces
class SoftAttentionFunction(torch.autograd.Function):
    """expected_soft_attention as its own operator
    (reference codebase/utils/monotonic_attention.py:79-152)."""

    @staticmethod
    def forward(ctx, alpha, soft_energy, padding_mask, chunk_size, eps):
        lib = _lib.load()
        dev = _lib.require_cuda(alpha, soft_energy, padding_mask)
        n, t, s = alpha.shape
        if tuple(soft_energy.shape) != (n, t, s):
            raise ValueError("soft_energy must have the shape of alpha")
        if s > _lib.SOFT_ATTENTION_MAX_SRC:
            raise ValueError(
                f"stand-alone expected_soft_attention keeps 6 fp32 rows on chip: src_len {s} exceeds "
                f"{_lib.SOFT_ATTENTION_MAX_SRC}; the fused operator (ops.mma_train) handles rows up to "
                f"{_lib.MMA_MAX_SRC}")
        a = alpha.contiguous()
        e = soft_energy.contiguous()
        mask = _mask_u8(padding_mask, n, s, dev)
        flags = _lib.MMA_ENERGY_F16_FILL if e.dtype == torch.float16 else 0
        beta = torch.empty_like(a)
        chunk = int(chunk_size) if chunk_size else 0
        with torch.cuda.device(dev):
            rc = lib.simulst_soft_attention_fwd(
                _lib.ptr(a), _lib.dtype_enum(a.dtype), _lib.ptr(e), _lib.dtype_enum(e.dtype),
                _lib.ptr(mask), _lib.ptr(beta), n, t, s, float(eps), chunk, flags,
                _lib.ptr(_lib.status_word(dev)), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_soft_attention_fwd")
        _lib.maybe_check(dev)
        ctx.save_for_backward(a, e, mask)
        ctx.cfg = (n, t, s, float(eps), chunk, flags)
        return beta

    @staticmethod
    def backward(ctx, g_beta):
        lib = _lib.load()
        a, e, mask = ctx.saved_tensors
        n, t, s, eps, chunk, flags = ctx.cfg
        dev = a.device
        gb = g_beta.contiguous().to(a.dtype)
        ga = torch.empty_like(a)
        ge = torch.empty_like(e)
        with torch.cuda.device(dev):
            rc = lib.simulst_soft_attention_bwd(
                _lib.ptr(a), _lib.dtype_enum(a.dtype), _lib.ptr(e), _lib.dtype_enum(e.dtype),
                _lib.ptr(mask), _lib.ptr(gb), _lib.ptr(ga), _lib.ptr(ge),
                n, t, s, eps, chunk, flags, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_soft_attention_bwd")
        return ga, ge, None, None, None


class MassPreservationFunction(torch.autograd.Function):
    """mass_preservation (reference codebase/utils/monotonic_attention.py:155-197).
    Like the reference it writes IN PLACE when there is no mask or with left padding
    (``alpha[:, :, -1] = residuals``); with right padding the reference builds new tensors
    (masked_fill / scatter_add), so the input is cloned first."""

    @staticmethod
    def forward(ctx, alpha, padding_mask, left_padding):
        lib = _lib.load()
        dev = _lib.require_cuda(alpha, padding_mask)
        n, t, s = alpha.shape
        mask = _mask_u8(padding_mask, n, s, dev)
        in_dtype = alpha.dtype
        in_place = (padding_mask is None) and alpha.dtype == torch.float32 and alpha.is_contiguous()
        work = alpha if in_place else alpha.float().contiguous().clone()
        side = torch.empty((n, t, 2), dtype=torch.float32, device=dev)
        flags = _lib.MMA_LEFT_PADDING if left_padding else 0
        with torch.cuda.device(dev):
            rc = lib.simulst_mass_preservation_fwd(
                _lib.ptr(work), _lib.ptr(mask), _lib.ptr(side), n, t, s, flags,
                _lib.ptr(_lib.status_word(dev)), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mass_preservation_fwd")
        _lib.maybe_check(dev)
        ctx.save_for_backward(mask, side)
        ctx.cfg = (n, t, s, flags, in_dtype)
        if in_place:
            ctx.mark_dirty(alpha)
            return alpha
        out = work.to(in_dtype)
        if padding_mask is None:           # reference mutates its argument in this branch
            alpha.copy_(out)
            ctx.mark_dirty(alpha)
            return alpha
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mask, side = ctx.saved_tensors
        n, t, s, flags, in_dtype = ctx.cfg
        dev = g.device
        gin = g.contiguous().float()
        gout = torch.empty_like(gin)
        with torch.cuda.device(dev):
            rc = lib.simulst_mass_preservation_bwd(
                _lib.ptr(gin), _lib.ptr(mask), _lib.ptr(side), _lib.ptr(gout), n, t, s, flags,
                _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mass_preservation_bwd")
        return gout.to(in_dtype), None, None


class MovingSumFunction(torch.autograd.Function):
    """moving_sum (functions.py:69-125).  The reference builds it from conv1d, so it is
    differentiable there; the adjoint of a sliding-window sum is the mirrored window:
    grad_x = moving_sum(grad_out, end_idx, start_idx)."""

    @staticmethod
    def forward(ctx, x, start_idx, end_idx):
        ctx.window = (int(start_idx), int(end_idx))
        return _moving_sum_raw(x, int(start_idx), int(end_idx))

    @staticmethod
    def backward(ctx, g):
        start_idx, end_idx = ctx.window
        return _moving_sum_raw(g, end_idx, start_idx), None, None


def _moving_sum_raw(x: Tensor, start_idx: int, end_idx: int) -> Tensor:
    lib = _lib.load()
    dev = _lib.require_cuda(x)
    n, t, s = x.shape
    xc = x.contiguous()
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        rc = lib.simulst_moving_sum(_lib.ptr(xc), _lib.ptr(out), _lib.dtype_enum(xc.dtype), n * t, s,
                                    start_idx, end_idx, _lib.stream_ptr(dev))
    _lib.check(rc, "simulst_moving_sum")
    return out


def moving_sum(x: Tensor, start_idx: int, end_idx: int) -> Tensor:
    """functions.py:69-125 over the last axis of a [N, T, S] tensor; differentiable like the
    reference's conv1d formulation."""
    assert start_idx > 0 and end_idx > 0
    return MovingSumFunction.apply(x, start_idx, end_idx)


class CumprodFunction(torch.autograd.Function):
    """exclusive_cumprod / safe_cumprod along the last axis (functions.py:20-66), differentiable
    like the reference's exp(cumsum(log(x + eps)))."""

    @staticmethod
    def forward(ctx, x, eps, inclusive, status):
        lib = _lib.load()
        dev = _lib.require_cuda(x)
        s = x.shape[-1]
        xc = x.contiguous()
        out = torch.empty_like(xc)
        rows = xc.numel() // max(s, 1)
        with torch.cuda.device(dev):
            rc = lib.simulst_exclusive_cumprod(_lib.ptr(xc), _lib.ptr(out), _lib.dtype_enum(xc.dtype), rows, s,
                                               float(eps), 1 if inclusive else 0,
                                               _lib.ptr(status), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_exclusive_cumprod")
        ctx.save_for_backward(xc, out)
        ctx.cfg = (float(eps), bool(inclusive))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        xc, out = ctx.saved_tensors
        eps, inclusive = ctx.cfg
        dev = xc.device
        s = xc.shape[-1]
        rows = xc.numel() // max(s, 1)
        gc = g.contiguous().to(xc.dtype)
        gx = torch.empty_like(xc)
        with torch.cuda.device(dev):
            rc = lib.simulst_cumprod_bwd(_lib.ptr(xc), _lib.ptr(out), _lib.ptr(gc), _lib.ptr(gx),
                                         _lib.dtype_enum(xc.dtype), rows, s, eps, 1 if inclusive else 0,
                                         _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_cumprod_bwd")
        return gx, None, None, None


def exclusive_cumprod_lastdim(x: Tensor, eps: float, inclusive: bool = False,
                              status: Optional[Tensor] = None) -> Tensor:
    """`status`: device word that receives SIMULST_ST_NEGPROD (default: the per-device word)."""
    if status is None:
        status = _lib.status_word(_lib.require_cuda(x))
    out = CumprodFunction.apply(x, eps, inclusive, status)
    _lib.maybe_check(x.device)
    return out


class PChooseFunction(torch.autograd.Function):
    """sigmoid(energy + noise) (reference codebase/utils/p_choose_strategy.py:56-76)."""

    @staticmethod
    def forward(ctx, energy, noise):
        lib = _lib.load()
        dev = _lib.require_cuda(energy, noise)
        e = energy.contiguous()
        nz = noise.contiguous().to(e.dtype) if noise is not None else None
        out = torch.empty_like(e)
        with torch.cuda.device(dev):
            rc = lib.simulst_p_choose(_lib.ptr(e), _lib.ptr(nz), _lib.ptr(out), _lib.dtype_enum(e.dtype),
                                      e.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_p_choose")
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        # d sigmoid = p (1 - p): elementwise on the producer side of the path (torch op)
        return g * p * (1 - p), None


# ----------------------------------------------------------------------------- incremental step
def mma_step(p_choose: Tensor, head_step: Tensor, soft_energy: Optional[Tensor] = None,
             src_lengths: Optional[Tensor] = None, mass_preservation: bool = True):
    """One decoding step for R = bsz*heads rows (reference
    modules/monotonic_multihead_attention.py:171-299).  `head_step` [R] int64 is updated in
    place.  Returns (head_read [R] bool, alpha [R,S], beta [R,S] or None)."""
    lib = _lib.load()
    dev = _lib.require_cuda(p_choose, head_step, soft_energy, src_lengths)
    r, s = p_choose.shape
    p = p_choose.contiguous()
    e = soft_energy.contiguous() if soft_energy is not None else None
    if head_step.dtype != torch.int64 or not head_step.is_contiguous() or head_step.numel() != r:
        raise ValueError("head_step must be a contiguous int64 tensor with bsz*heads elements")
    lens = src_lengths.to(torch.int32).contiguous() if src_lengths is not None else None
    head_read = torch.empty(r, dtype=torch.uint8, device=dev)
    alpha = torch.empty_like(p)
    beta = torch.empty_like(e) if e is not None else None
    flags = _lib.MMA_MASS_PRESERVATION if mass_preservation else 0
    with torch.cuda.device(dev):
        rc = lib.simulst_mma_step(_lib.ptr(p), _lib.dtype_enum(p.dtype), _lib.ptr(e),
                                  _lib.dtype_enum(e.dtype) if e is not None else 0, _lib.ptr(lens),
                                  _lib.ptr(head_step), _lib.ptr(head_read), _lib.ptr(alpha),
                                  _lib.ptr(beta), r, s, flags, _lib.stream_ptr(dev))
    _lib.check(rc, "simulst_mma_step")
    return head_read.view(torch.bool), alpha, beta


# ----------------------------------------------------------------------------- CIF
class CIFFunction(torch.autograd.Function):
    """cif_function (reference codebase/models/torch_cif/cif.py:23-196): plan + gather forward,
    frame-gather + row-scan backward."""

    @staticmethod
    def forward(ctx, input, alpha, padding_mask, target_lengths, beta, tail_thres, eps):
        lib = _lib.load()
        training = target_lengths is not None
        host_lengths = training and not target_lengths.is_cuda
        dev = _lib.require_cuda(input, alpha, padding_mask, None if host_lengths else target_lengths)
        b, s, c = input.shape
        x = input.contiguous()
        a = alpha.contiguous()
        mask = _mask_u8(padding_mask, b, s, dev)
        status = _lib.status_word(dev)
        # (separate allocations on purpose: slicing one packed buffer costs more dispatcher time than the
        #  caching allocator does -- measured 0.32 -> 0.37 ms per call)
        csum = torch.empty((b, s), dtype=torch.float32, device=dev)
        scale = torch.empty(b, dtype=torch.float32, device=dev)
        alpha_sum = torch.empty(b, dtype=torch.float32, device=dev)
        lengths0 = torch.empty(b, dtype=torch.int64, device=dev)
        counters = torch.zeros(2, dtype=torch.int32, device=dev)      # t_max, t_max2
        desired = tl = None
        if training:
            if host_lengths:
                # extension over the reference (whose target_lengths live on the input's device): lengths
                # that are still on the host give T without the device read of cif.py:72 -- no sync at all
                t_cap = int(target_lengths.max()) if b > 0 else 0
                target_lengths = target_lengths.to(dev, non_blocking=True)
            tl = target_lengths.long().contiguous()
            # cif.py:68 -- evaluated in the INPUT dtype, as the reference does
            desired = (beta * target_lengths.type_as(x) + eps).float().contiguous()
            if not host_lengths:
                t_cap = int(tl.max()) if b > 0 else 0                  # host read, cif.py:72
        # segment table rows: slots 0..T+1 (training) / up to floor(S/beta)+1 fires (inference)
        seg_stride = int(lib.simulst_cif_seg_stride(1 if training else 0, t_cap if training else 0, s, float(beta)))
        seg_first = torch.empty((b, seg_stride), dtype=torch.int32, device=dev)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_cif_plan(_lib.ptr(a), _lib.dtype_enum(a.dtype), _lib.ptr(mask),
                                      _lib.ptr(desired), _lib.ptr(tl), _lib.ptr(csum), _lib.ptr(scale),
                                      _lib.ptr(alpha_sum), _lib.ptr(lengths0),
                                      counters.data_ptr(), _lib.ptr(seg_first), seg_stride,
                                      b, s, float(beta), _lib.ptr(status), st)
            _lib.check(rc, "simulst_cif_plan")
            if not training:
                t_cap = int(counters[0].item()) if b > 0 else 0        # host read, cif.py:76
            t_alloc = t_cap if training else t_cap + 1
            out = torch.empty((b, t_alloc, c), dtype=x.dtype, device=dev)
            delays = torch.empty((b, t_alloc), dtype=x.dtype, device=dev)
            tail_w = lengths1 = None
            if not training:
                tail_w = torch.empty(b, dtype=torch.float32, device=dev)
                lengths1 = torch.empty(b, dtype=torch.int64, device=dev)
            rc = lib.simulst_cif_fwd(_lib.ptr(x), _lib.dtype_enum(x.dtype), _lib.ptr(csum), _lib.ptr(scale),
                                     _lib.ptr(a), _lib.dtype_enum(a.dtype), _lib.ptr(mask),
                                     _lib.ptr(seg_first), seg_stride,
                                     _lib.ptr(out), _lib.ptr(delays), _lib.ptr(tail_w),
                                     _lib.ptr(lengths0), _lib.ptr(lengths1),
                                     counters.data_ptr() + 4, b, s, c, t_cap, t_alloc,
                                     float(beta), float(tail_thres), 1 if training else 0, st)
            _lib.check(rc, "simulst_cif_fwd")
        _lib.maybe_check(dev)
        if training:
            t_out = t_cap
            lengths = lengths0
        else:
            t_out = int(counters[1].item()) if b > 0 else 0            # host read, cif.py:181
            lengths = lengths1
        ctx.save_for_backward(x, a, mask, csum, scale, alpha_sum, tail_w, lengths0, lengths1)
        ctx.cfg = (b, s, c, t_cap, t_out, float(beta), float(tail_thres), training, alpha.dtype)
        out_v = out[:, :t_out]
        delays_v = delays[:, :t_out]
        asum = alpha_sum.to(alpha.dtype)
        tail_ret = tail_w if tail_w is not None else alpha_sum.new_empty(0)
        ctx.mark_non_differentiable(lengths, tail_ret)
        return out_v, delays_v, asum, lengths, tail_ret

    @staticmethod
    def backward(ctx, g_out, g_delays, g_asum, _g_len, _g_tail):
        lib = _lib.load()
        x, a, mask, csum, scale, alpha_sum, tail_w, lengths0, lengths1 = ctx.saved_tensors
        b, s, c, t_cap, t_out, beta, tail_thres, training, a_dtype = ctx.cfg
        dev = x.device
        go = (g_out.contiguous().to(x.dtype) if g_out is not None
              else torch.zeros((b, t_out, c), dtype=x.dtype, device=dev))
        gd = g_delays.contiguous().to(x.dtype) if g_delays is not None else None
        gs = g_asum.contiguous().float() if g_asum is not None else None
        gx = torch.empty_like(x)
        ga = torch.empty_like(a)
        ws = torch.empty(int(lib.simulst_cif_workspace_bytes(b, s)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_cif_bwd(_lib.ptr(x), _lib.dtype_enum(x.dtype), _lib.ptr(csum), _lib.ptr(scale),
                                     _lib.ptr(a), _lib.dtype_enum(a.dtype), _lib.ptr(mask),
                                     _lib.ptr(go), _lib.ptr(gd), _lib.ptr(tail_w), _lib.ptr(lengths0),
                                     _lib.ptr(lengths1), _lib.ptr(alpha_sum), _lib.ptr(gs),
                                     _lib.ptr(gx), _lib.ptr(ga), _lib.ptr(ws),
                                     b, s, c, t_cap, t_out, beta, tail_thres, 1 if training else 0,
                                     _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_cif_bwd")
        return gx, ga, None, None, None, None, None


# ----------------------------------------------------------------------------- latency loss
class DALFunction(torch.autograd.Function):
    """DifferentiableAverageLagging(delays, src_lens, ref_lens, target_padding_mask) as the
    reference's criteria call it (codebase/criterion/mma_criterion.py:172-177,
    codebase/criterion/cif_criterion.py:211-216).  Returns [N, 1] like SimulEval's function."""

    @staticmethod
    def forward(ctx, delays, src_lens, ref_lens, target_padding_mask):
        lib = _lib.load()
        dev = _lib.require_cuda(delays, src_lens, ref_lens, target_padding_mask)
        if delays.dim() != 2:
            raise ValueError("delays must be [N, tgt_len]")
        n, t = delays.shape
        d = delays.contiguous().float()
        src = src_lens.reshape(n).long().contiguous()
        ref = ref_lens.reshape(n).long().contiguous() if ref_lens is not None else None
        mask = _mask_u8(target_padding_mask, n, t, dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_dal_fwd(_lib.ptr(d), _lib.ptr(src), _lib.ptr(ref), _lib.ptr(mask), _lib.ptr(out),
                                     n, t, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_dal_fwd")
        ctx.save_for_backward(d, src, ref, mask)
        ctx.in_dtype = delays.dtype
        return out.view(n, 1).to(delays.dtype)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        d, src, ref, mask = ctx.saved_tensors
        n, t = d.shape
        dev = d.device
        gd = g.reshape(n).contiguous().float()
        out = torch.empty_like(d)
        with torch.cuda.device(dev):
            rc = lib.simulst_dal_bwd(_lib.ptr(d), _lib.ptr(src), _lib.ptr(ref), _lib.ptr(mask), _lib.ptr(gd),
                                     _lib.ptr(out), n, t, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_dal_bwd")
        return out.to(ctx.in_dtype), None, None, None


def differentiable_average_lagging(delays: Tensor, src_lens: Tensor, ref_lens: Optional[Tensor] = None,
                                   target_padding_mask: Optional[Tensor] = None) -> Tensor:
    """Same call shape as SimulEval's ``DifferentiableAverageLagging`` (the entry of
    ``LATENCY_METRICS`` the reference's criteria use); one kernel launch forward, one backward."""
    return DALFunction.apply(delays, src_lens, ref_lens, target_padding_mask)


# ----------------------------------------------------------------------------- SSNT lattice loss
class SSNTFunction(torch.autograd.Function):
    """ssnt_loss / ssnt_loss_mem (reference codebase/criterion/ssnt_loss/ssnt_loss.py:45-271) up
    to the reduction: (log_probs, emit) -> (loss [N], lattice, log_p_choose).  One forward and one
    backward launch; the backward recomputes the scans from the saved lattice."""

    @staticmethod
    def forward(ctx, log_probs, emit, targets, source_lengths, target_lengths, emit_is_logits,
                neg_inf, fastemit_lambda, flat):
        lib = _lib.load()
        dev = _lib.require_cuda(log_probs, emit, targets, source_lengths, target_lengths)
        n = source_lengths.numel()
        lp = log_probs.contiguous()
        em = emit.contiguous()
        tg = targets.contiguous().long()
        src = source_lengths.contiguous().long()
        tgt = target_lengths.contiguous().long()
        if flat:
            rows, s = em.shape
            t = 0
            row_off = (torch.cumsum(tgt, 0) - tgt).contiguous()
            lat_off = (torch.cumsum(tgt + 1, 0) - (tgt + 1)).contiguous()
            lattice = torch.empty((rows + n, s), dtype=torch.float32, device=dev)
        else:
            n_e, t, s = em.shape
            if n_e != n:
                raise ValueError(f"emit has batch {n_e}, lengths have {n}")
            rows = n * t
            row_off = lat_off = None
            lattice = torch.empty((n, t, s), dtype=torch.float32, device=dev)
        v = lp.shape[-1]
        if lp.numel() != rows * s * v or tg.numel() != rows:
            raise ValueError("log_probs / targets do not match the emission tensor")
        if s > _lib.SSNT_MAX_SRC:
            raise ValueError(f"src_len {s} exceeds the on-chip row limit {_lib.SSNT_MAX_SRC}")
        log_p = torch.empty(em.shape, dtype=torch.float32, device=dev)
        loss = torch.empty(n, dtype=torch.float32, device=dev)
        status = _lib.status_word(dev)
        with torch.cuda.device(dev):
            rc = lib.simulst_ssnt_fwd(_lib.ptr(lp), _lib.dtype_enum(lp.dtype), _lib.ptr(tg), _lib.ptr(em),
                                      _lib.dtype_enum(em.dtype), 1 if emit_is_logits else 0, _lib.ptr(src),
                                      _lib.ptr(tgt), _lib.ptr(row_off), _lib.ptr(lat_off), _lib.ptr(lattice),
                                      _lib.ptr(log_p), _lib.ptr(loss), n, t, s, v, float(neg_inf),
                                      float(fastemit_lambda), _lib.ptr(status), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_ssnt_fwd")
        ctx.save_for_backward(lp, em, tg, src, tgt, row_off, lat_off, lattice)
        ctx.cfg = (n, t, s, v, bool(emit_is_logits), float(neg_inf), float(fastemit_lambda))
        ctx.set_materialize_grads(False)
        return loss, lattice, log_p

    @staticmethod
    def backward(ctx, g_loss, g_lattice, g_log_p):
        lib = _lib.load()
        lp, em, tg, src, tgt, row_off, lat_off, lattice = ctx.saved_tensors
        n, t, s, v, is_logits, neg_inf, lam = ctx.cfg
        dev = em.device
        gl = (g_loss.contiguous().float() if g_loss is not None
              else torch.zeros(n, dtype=torch.float32, device=dev))
        glat = g_lattice.contiguous().float() if g_lattice is not None else None
        glp = g_log_p.contiguous().float() if g_log_p is not None else None
        g_emit = torch.empty_like(em)
        # the gather's adjoint is a scatter into zeros, as in autograd; only when someone wants it
        g_probs = torch.zeros_like(lp) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(dev):
            rc = lib.simulst_ssnt_bwd(_lib.ptr(lp), _lib.dtype_enum(lp.dtype), _lib.ptr(tg), _lib.ptr(em),
                                      _lib.dtype_enum(em.dtype), 1 if is_logits else 0, _lib.ptr(src),
                                      _lib.ptr(tgt), _lib.ptr(row_off), _lib.ptr(lat_off), _lib.ptr(lattice),
                                      _lib.ptr(gl), _lib.ptr(glat), _lib.ptr(glp), _lib.ptr(g_emit),
                                      _lib.ptr(g_probs), n, t, s, v, neg_inf, lam, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_ssnt_bwd")
        return g_probs, g_emit, None, None, None, None, None, None, None


def logprob_check(log_probs: Tensor, neg_inf: float = -1e8):
    """prob_check(tensor, neg_inf=..., logp=True) of the reference (ssnt_loss.py:29-42) as one
    streaming pass into the device status word (raise via check_status / strict mode)."""
    lib = _lib.load()
    dev = _lib.require_cuda(log_probs)
    x = log_probs.contiguous()
    with torch.cuda.device(dev):
        rc = lib.simulst_logprob_check(_lib.ptr(x), _lib.dtype_enum(x.dtype), x.numel(), float(neg_inf),
                                       _lib.ptr(_lib.status_word(dev)), _lib.stream_ptr(dev))
    _lib.check(rc, "simulst_logprob_check")
    _lib.maybe_check(dev)


# ----------------------------------------------------------------------------- CTC best alignment
def ctc_best_alignment(log_prob: Tensor, targets: Tensor, input_lengths: Tensor, target_lengths: Tensor,
                       blank: int = 0, as_labels: bool = False, max_target_length: Optional[int] = None,
                       return_nll: bool = False):
    """Viterbi forward + back-trace in one launch (reference
    codebase/criterion/best_alignment/{best_alignment.cu:58-202, __init__.py:25-111}).
    log_prob (S, N, V); targets (N, T); lengths (N,) on the device.  `max_target_length`: the
    width of the state row (2*max+1); defaults to targets.size(1) -- no host read of the lengths
    (the reference copies both length vectors to the host on every call)."""
    lib = _lib.load()
    dev = _lib.require_cuda(log_prob, targets, input_lengths, target_lengths)
    if log_prob.dim() != 3 or targets.dim() != 2:
        raise ValueError("log_prob must be (S, N, V) and targets (N, T)")
    s, n, v = log_prob.shape
    lp = log_prob.contiguous()
    tg = targets.contiguous().long()
    il = input_lengths.contiguous().long()
    tl = target_lengths.contiguous().long()
    t_max = int(max_target_length) if max_target_length is not None else tg.shape[1]
    if t_max > tg.shape[1]:
        raise ValueError("max_target_length exceeds targets.size(1)")
    ws = torch.empty(max(int(lib.simulst_ctc_workspace_bytes(n, s, t_max)), 1), dtype=torch.uint8, device=dev)
    states = torch.empty((n, s), dtype=torch.int64, device=dev)
    labels = torch.empty((n, s), dtype=torch.int64, device=dev) if as_labels else None
    nll = torch.empty(n, dtype=torch.float32, device=dev) if return_nll else None
    with torch.cuda.device(dev):
        rc = lib.simulst_ctc_best_alignment(_lib.ptr(lp), _lib.dtype_enum(lp.dtype), _lib.ptr(tg), tg.shape[1],
                                            _lib.ptr(il), _lib.ptr(tl), int(blank), _lib.ptr(ws), _lib.ptr(nll),
                                            _lib.ptr(states), _lib.ptr(labels), n, s, v, t_max,
                                            _lib.stream_ptr(dev))
    _lib.check(rc, "simulst_ctc_best_alignment")
    out = labels if as_labels else states
    return (out, nll) if return_nll else out
