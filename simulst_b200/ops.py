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


class MMATrainFunction(torch.autograd.Function):
    """(p_choose, soft_energy) -> (alpha, beta): steps 2-3 of
    MonotonicAttention.monotonic_attention_process_train
    (reference codebase/modules/monotonic_multihead_attention.py:318-347) in one launch;
    backward is one launch that recomputes every scan."""

    @staticmethod
    def forward(ctx, p_choose: Tensor, soft_energy: Optional[Tensor], padding_mask: Optional[Tensor],
                eps: float, mass_preservation: bool, chunk_size: Optional[int],
                left_padding: bool = False):
        lib = _lib.load()
        dev = _lib.require_cuda(p_choose, soft_energy, padding_mask)
        if p_choose.dim() != 3:
            raise ValueError("p_choose must be [bsz*heads, tgt_len, src_len]")
        n, t, s = p_choose.shape
        if s > _lib.MMA_MAX_SRC:
            raise ValueError(f"src_len {s} exceeds the on-chip row limit {_lib.MMA_MAX_SRC}")
        soft = soft_energy is not None
        p = p_choose.contiguous()
        e = None
        flags = 0
        if soft:
            if tuple(soft_energy.shape) != (n, t, s):
                raise ValueError("soft_energy must have the shape of p_choose")
            if soft_energy.dtype == torch.float16:
                flags |= _lib.MMA_ENERGY_F16_FILL      # -1e4 fill, monotonic_attention.py:106
            e = soft_energy.contiguous()
            if e.dtype != p.dtype:                      # kernel wants one activation dtype
                common = torch.promote_types(e.dtype, p.dtype)
                e, p = e.to(common), p.to(common)
            flags |= _lib.MMA_SOFT
        if mass_preservation:
            flags |= _lib.MMA_MASS_PRESERVATION
        if left_padding:
            flags |= _lib.MMA_LEFT_PADDING
        mask = _mask_u8(padding_mask, n, s, dev)
        alpha = torch.empty((n, t, s), dtype=torch.float32, device=dev)
        beta = torch.empty((n, t, s), dtype=torch.float32, device=dev) if soft else None
        side = torch.empty((n, t, 2), dtype=torch.float32, device=dev) if mass_preservation else None
        status = _lib.status_word(dev)
        chunk = int(chunk_size) if chunk_size else 0
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_fwd(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, _lib.ptr(mask),
                _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(side),
                n, t, s, float(eps), chunk, flags, _lib.ptr(status), _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_fwd")
        _lib.maybe_check(dev)
        ctx.save_for_backward(p, e, mask, alpha, side)
        ctx.cfg = (n, t, s, float(eps), chunk, flags, soft)
        ctx.in_dtypes = (p_choose.dtype, soft_energy.dtype if soft else None)
        ctx.mark_non_differentiable()
        if soft:
            return alpha, beta
        return alpha, alpha.new_empty(0)

    @staticmethod
    def backward(ctx, g_alpha, g_beta):
        lib = _lib.load()
        p, e, mask, alpha, side = ctx.saved_tensors
        n, t, s, eps, chunk, flags, soft = ctx.cfg
        dev = p.device
        ga = g_alpha.contiguous().float() if g_alpha is not None else None
        gb = g_beta.contiguous().float() if (soft and g_beta is not None) else None
        grad_p = torch.empty_like(p)
        grad_e = torch.empty_like(e) if soft else None
        with torch.cuda.device(dev):
            rc = lib.simulst_mma_train_bwd(
                _lib.ptr(p), _lib.dtype_enum(p.dtype), _lib.ptr(e),
                _lib.dtype_enum(e.dtype) if soft else 0, _lib.ptr(mask),
                _lib.ptr(alpha), _lib.ptr(side), _lib.ptr(ga), _lib.ptr(gb),
                _lib.ptr(grad_p), _lib.dtype_enum(p.dtype), _lib.ptr(grad_e),
                _lib.dtype_enum(e.dtype) if soft else 0,
                n, t, s, eps, chunk, flags, _lib.stream_ptr(dev))
        _lib.check(rc, "simulst_mma_train_bwd")
        p_dt, e_dt = ctx.in_dtypes
        if grad_p.dtype != p_dt:
            grad_p = grad_p.to(p_dt)
        if soft and grad_e.dtype != e_dt:
            grad_e = grad_e.to(e_dt)
        return grad_p, grad_e, None, None, None, None, None


def mma_train(p_choose: Tensor, soft_energy: Optional[Tensor] = None,
              padding_mask: Optional[Tensor] = None, eps: float = 1e-6,
              mass_preservation: bool = True, chunk_size: Optional[int] = None,
              left_padding: bool = False):
    """Fused expected alignment (+ mass preservation) (+ expected soft attention).
    Returns (alpha [N,T,S] fp32, beta [N,T,S] fp32); beta is alpha for hard attention."""
    alpha, beta = MMATrainFunction.apply(p_choose, soft_energy, padding_mask, eps,
                                         mass_preservation, chunk_size, left_padding)
    if soft_energy is None:
        beta = alpha
    return alpha, beta
