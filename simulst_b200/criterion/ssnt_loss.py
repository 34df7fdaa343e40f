"""Mirror of ``codebase/criterion/ssnt_loss/ssnt_loss.py`` of the reference (submodule ssnt_loss
@ a5af91e): same function names, arguments, defaults and return triples; the lattice recurrence
and its gradient run in the sm_100a kernels (``simulst_ssnt_fwd/bwd``), CUDA tensors only.

Range checks: the reference asserts on the host three times per call (``prob_check`` on the whole
``[.., S, V]`` log-prob tensor, on ``emit_probs`` and on the result).  Here the same conditions are
recorded in the device status word -- one streaming pass over ``log_probs``
(``simulst_logprob_check``), the emission check inside the lattice kernel -- and raised as the same
``AssertionError`` by ``simulst_b200.check_status()`` (eagerly under ``set_strict(True)``).
"""
from typing import Optional

import torch
from torch import Tensor

from .. import _lib, ops


def lengths_to_padding_mask(lens):
    """ssnt_loss.py:6-10."""
    bsz, max_lens = lens.size(0), torch.max(lens).item()
    mask = torch.arange(max_lens, device=lens.device).view(1, max_lens)
    return mask.expand(bsz, -1) >= lens.view(bsz, 1).expand(-1, max_lens)


def exclusive_cumsum(tensor, dim: int):
    """ssnt_loss.py:22-26: [0, x1, x1+x2, ...] along `dim` (torch ops: helper used by callers and
    by the reference's own test, not by the kernel path)."""
    shifted = tensor.roll(1, dims=dim)
    shifted.select(dim, 0).fill_(0)
    return shifted.cumsum(dim)


def log_exclusive_cumprod(tensor, dim: int):
    """ssnt_loss.py:14-19: exclusive cumprod of a tensor given in log space."""
    return exclusive_cumsum(tensor, dim)


def prob_check(tensor, eps=1e-10, neg_inf=-1e8, logp=False):
    """ssnt_loss.py:29-42 (literal, host-syncing form for callers that want it)."""
    assert not torch.isnan(tensor).any(), (
        "Nan in a probability tensor."
    )
    if logp:
        assert tensor.le(0).all() and tensor.ge(neg_inf).all(), (
            "Incorrect values in a log-probability tensor"
            ", -inf <= tensor <= 0"
        )
    else:
        assert tensor.le(1.0 + eps).all() and tensor.ge(0.0 - eps).all(), (
            "Incorrect values in a probability tensor"
            ", 0.0 <= tensor <= 1.0"
        )


def _run(log_probs, targets, source_lengths, target_lengths, emit_logits, emit_probs, neg_inf, reduction,
         fastemit_lambda, flat):
    ops.logprob_check(log_probs, neg_inf)
    if emit_logits is None:
        assert emit_probs is not None, "emit_probs and emit_logits cannot both be None."
        emit, is_logits = emit_probs, False
    else:
        emit, is_logits = emit_logits, True
    loss, lattice, log_p_choose = ops.SSNTFunction.apply(
        log_probs, emit, targets, source_lengths, target_lengths, is_logits, float(neg_inf),
        float(fastemit_lambda), flat)
    if reduction == "sum":
        loss = loss.sum()
    elif reduction == "mean":
        loss = loss.mean()
    _lib.maybe_check(emit.device)
    return loss, lattice, log_p_choose


def ssnt_loss(
    log_probs: Tensor,
    targets: Tensor,
    source_lengths: Tensor,
    target_lengths: Tensor,
    emit_logits: Optional[Tensor] = None,
    emit_probs: Optional[Tensor] = None,
    neg_inf: float = -1e4,
    reduction="none",
    fastemit_lambda=0
):
    """ssnt_loss.py:45-151.  log_probs (N, T, S, V), targets (N, T), emit (N, T, S).
    Returns (-log_alpha at the sequence ends [reduced], lattice (N, T, S), log_p_choose (N, T, S))."""
    return _run(log_probs, targets, source_lengths, target_lengths, emit_logits, emit_probs, neg_inf,
                reduction, fastemit_lambda, False)


def ssnt_loss_mem(
    log_probs: Tensor,
    targets: Tensor,
    source_lengths: Tensor,
    target_lengths: Tensor,
    emit_logits: Optional[Tensor] = None,
    emit_probs: Optional[Tensor] = None,
    neg_inf: float = -1e4,
    reduction="none",
    fastemit_lambda=0
):
    """ssnt_loss.py:154-271: targets concatenated over the batch -- log_probs (T_flat, S, V),
    targets (T_flat,), emit (T_flat, S); the lattice is the reference's (T_flat + N, S) buffer
    (each sample's rows start with its alpha_0 row)."""
    return _run(log_probs, targets, source_lengths, target_lengths, emit_logits, emit_probs, neg_inf,
                reduction, fastemit_lambda, True)
