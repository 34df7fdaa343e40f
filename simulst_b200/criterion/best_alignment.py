"""Mirror of ``codebase/criterion/best_alignment/__init__.py`` of the reference: the same
``best_alignment(log_prob, targets, input_lengths, target_lengths, blank=0, as_labels=False)``,
served by ONE sm_100a kernel (Viterbi forward in shared memory, byte-sized jump table, device
back-trace) instead of the JIT-built ATen extension plus an S-iteration Python loop.

One upstream defect is not reproduced: with ``as_labels=True`` the reference gathers
``targets[states // 2]`` for EVERY frame before selecting with ``torch.where`` (:101-106), so a
path that reaches the final blank of the longest target indexes one past the end of ``targets``
(an exception on CPU, a device-side assert on CUDA).  The label at such a frame is ``blank`` by
the ``where``; the kernel writes that directly.
"""
import torch

from .. import ops


def best_alignment(
    log_prob: torch.Tensor,
    targets: torch.Tensor,
    input_lengths: torch.Tensor,
    target_lengths: torch.Tensor,
    blank: int = 0,
    as_labels: bool = False
):
    """Get best alignment (maximum probability sequence of ctc states) conditioned on log
    probabilities and target sequences.  log_prob (S, N, V) after log_softmax; targets (N, T);
    returns (N, S) int64 states in [0, 2T+1), or labels in [0, V) with ``as_labels``."""
    return ops.ctc_best_alignment(log_prob, targets, input_lengths, target_lengths, blank, as_labels)


def quantity_targets(states: torch.Tensor, encoder_padding_mask=None):
    """What CIFCriterion derives from the state sequence (codebase/criterion/cif_criterion.py:
    249-262): ``boundary`` = last frame of every non-blank segment, and the cumulative count of
    boundaries (the target of the quantity loss).  torch ops on an (N, S) integer tensor."""
    seg_ids = states.div(2, rounding_mode='floor')
    boundary = (seg_ids != seg_ids.roll(-1, dims=1)) & (states % 2 != 0)
    if encoder_padding_mask is not None:
        boundary = boundary.masked_fill(encoder_padding_mask, False)
    return boundary, boundary.cumsum(1)
