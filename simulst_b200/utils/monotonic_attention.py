"""Mirror of ``codebase/utils/monotonic_attention.py`` of the reference: identical
signatures, defaults, output dtypes and in-place behaviour; the math runs in the sm_100a
kernels.  ``mma_process_train`` is the fused form used by the attention module."""
from typing import Optional

import torch
from torch import Tensor

from .. import ops


def expected_alignment_from_p_choose(
    p_choose: Tensor,
    padding_mask: Optional[Tensor] = None,
    eps: float = 1e-6
):
    """monotonic_attention.py:12-76.  p_choose: bsz, tgt_len, src_len; returns alpha in
    p_choose's dtype (:72)."""
    alpha, _ = ops.mma_train(p_choose, None, padding_mask, eps=eps, mass_preservation=False)
    # The operator keeps its fp32 output for the recompute-based backward.  Callers of this
    # stand-alone function may write into the result (mass_preservation does, in place, like the
    # reference), so they get their own tensor.  The fused path (mma_process_train) has no copy.
    if alpha.dtype != p_choose.dtype:
        return alpha.type(p_choose.dtype)
    return alpha.clone()


def expected_soft_attention(
    alpha: Tensor,
    soft_energy: Tensor,
    padding_mask: Optional[Tensor] = None,
    chunk_size: Optional[int] = None,
    eps: float = 1e-10
):
    """monotonic_attention.py:79-152.  Returns beta in alpha's dtype, clamped to [0, 1]."""
    return ops.SoftAttentionFunction.apply(alpha, soft_energy, padding_mask, chunk_size, eps)


def mass_preservation(
    alpha: Tensor,
    padding_mask: Optional[Tensor] = None,
    left_padding: bool = False
):
    """monotonic_attention.py:155-197 (mutates `alpha` when there is no mask, like the
    reference's ``alpha[:, :, -1] = residuals``)."""
    if padding_mask is not None and not left_padding:
        assert not padding_mask[:, 0].any(), (
            "Find padding on the beginning of the sequence."
        )
    return ops.MassPreservationFunction.apply(alpha, padding_mask, left_padding)


def mma_process_train(p_choose: Tensor, soft_energy: Optional[Tensor],
                      padding_mask: Optional[Tensor] = None, eps: float = 1e-6,
                      mass_preservation: bool = True, chunk_size: Optional[int] = None):
    """alpha, beta of monotonic_attention_process_train in ONE forward launch
    (expected alignment -> mass preservation -> expected soft attention)."""
    return ops.mma_train(p_choose, soft_energy, padding_mask, eps=eps,
                         mass_preservation=mass_preservation, chunk_size=chunk_size)


def mma_process_train_with_delays(p_choose: Tensor, soft_energy: Optional[Tensor],
                                  padding_mask: Optional[Tensor] = None, eps: float = 1e-6,
                                  mass_preservation: bool = True, chunk_size: Optional[int] = None):
    """mma_process_train plus the expected delays ``sum_j (j+1) * alpha[n,i,j]`` ([N,T]) that
    MMACriterion.compute_latency_loss derives from alpha (reference
    codebase/criterion/mma_criterion.py:146-157), produced by the same launch."""
    return ops.mma_train_with_delays(p_choose, soft_energy, padding_mask, eps=eps,
                                     mass_preservation=mass_preservation, chunk_size=chunk_size)
