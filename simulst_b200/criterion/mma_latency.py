"""Latency loss of ``MMACriterion`` on the B200 path (reference
``codebase/criterion/mma_criterion.py:138-207``).

The reference re-reads every layer's ``alpha [bsz, heads, T, S]`` to form the expected delays
``sum_j (j+1) * alpha`` (:146-157) and then runs SimulEval's DifferentiableAverageLagging as T
rounds of cat/max (:172-177).  Here the delays come out of the alignment kernel itself
(``simulst_mma_train_fwd_delays``: the attention mixin keeps them on the module when
``with_expected_delays`` is set), DAL is one kernel (``simulst_dal_fwd/bwd``), and what is left --
the gather over layers*heads, the clip and the variance term (:179-200) -- are a handful of
torch ops on ``[bsz, layers*heads(, T)]`` tensors.  ``compute_latency_loss`` keeps the reference
method's name, arguments and return triple, so a criterion subclass can use it as its method.
"""
from typing import List, Optional

import torch
from torch import Tensor

from .. import ops

LATENCY_METRICS = {"differentiable_average_lagging": ops.differentiable_average_lagging}


def expected_delays_from_alpha(alpha_all: Tensor) -> Tensor:
    """mma_criterion.py:146-157 for callers that only hold alpha (no fused epilogue): one pass."""
    src_len = alpha_all.size(-1)
    steps = torch.arange(1, 1 + src_len, device=alpha_all.device).type_as(alpha_all)
    return torch.matmul(alpha_all, steps)


def latency_loss_from_delays(self, expected_delays: Tensor, bsz: int, heads_total: int, target: Tensor,
                             input_lengths: Tensor, encoder_padding_mask: Tensor):
    """mma_criterion.py:159-207 given ``expected_delays [bsz*layers*heads, T]`` (rows ordered
    utterance-major, like ``torch.cat(alpha_list, dim=1).view(-1, T, S)``)."""
    tgt_len = expected_delays.size(1)
    target_padding_mask = target == self.padding_idx
    target_lengths = (~target_padding_mask).sum(1)
    if isinstance(encoder_padding_mask, list):
        encoder_padding_mask = encoder_padding_mask[0]
    assert not encoder_padding_mask[:, 0].any(), "Only right padding is supported."
    encoder_lengths = (~encoder_padding_mask).sum(-1)

    def expand(t):
        return torch.repeat_interleave(t, heads_total, 0)

    metric = LATENCY_METRICS[self.latency_avg_type]
    expected_latency = metric(
        expected_delays,
        expand(encoder_lengths),
        expand(target_lengths),
        target_padding_mask=expand(target_padding_mask),
    )
    expected_latency = expected_latency.view(bsz, -1)
    if self.latency_gather_method == "average":
        expected_latency = expected_delays.mean(dim=1)      # as upstream (:184)
    elif self.latency_gather_method == "weighted_average":
        weights = torch.nn.functional.softmax(expected_latency, dim=1)
        expected_latency = torch.sum(expected_latency * weights, dim=1)
    elif self.latency_gather_method == "max":
        expected_latency = expected_latency.max(dim=1)[0]
    else:
        raise NotImplementedError
    avg_loss = self.latency_avg_weight * expected_latency.clip(min=0).sum()
    expected_delays_var = expected_delays.view(bsz, -1, tgt_len).var(dim=1).mean(dim=1).sum()
    latency_loss = avg_loss + self.latency_var_weight * expected_delays_var
    expected_latency = expected_latency * (input_lengths / encoder_lengths * self.ms_per_frame_shift)
    return latency_loss, expected_latency.sum(), expected_delays_var


def compute_latency_loss(self, model, sample, net_output, expected_delays: Optional[List[Tensor]] = None):
    """Drop-in for ``MMACriterion.compute_latency_loss(self, model, sample, net_output)``.

    ``expected_delays``: optional list, one ``[bsz*heads, T]`` tensor per decoder layer, as left on
    each attention module by the mixin (``module.expected_delays``); without it the delays are
    reduced from ``net_output[1]["attn_list"][l]["alpha"]`` like the reference does."""
    alpha_list = [item["alpha"] for item in net_output[1]["attn_list"]]
    num_layers = len(alpha_list)
    bsz, num_heads, tgt_len, src_len = alpha_list[0].size()
    if expected_delays is not None:
        assert len(expected_delays) == num_layers
        per_layer = [d.view(bsz, num_heads, tgt_len) for d in expected_delays]
    else:
        per_layer = [expected_delays_from_alpha(a) for a in alpha_list]
    # same row order as torch.cat(alpha_list, dim=1).view(-1, tgt_len, src_len)
    delays = torch.cat(per_layer, dim=1).reshape(-1, tgt_len)
    return latency_loss_from_delays(self, delays, bsz, num_layers * num_heads, sample["target"],
                                    sample["net_input"]["src_lengths"], net_output[-1]["encoder_padding_mask"])
