"""Oracle (CPU PyTorch) for the latency loss of the MMA criterion -- SURVEY 8f rank 1.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Reference: ``codebase/criterion/mma_criterion.py:138-207`` (``MMACriterion.compute_latency_loss``):
  :146-157  expected delays  ``sum_j (j+1) * alpha[n,i,j]``
  :159-177  lengths, ``LATENCY_METRICS[latency_avg_type](expected_delays, enc_len, tgt_len, mask)``
  :179-191  gather over layers*heads (``average`` / ``weighted_average`` / ``max``)
  :193      ``latency_avg_weight * clip(min=0).sum()``
  :195-200  variance of the expected delays over layers*heads
  :205-206  renormalisation to ms

Third-party arithmetic outside /root/reference: ``DifferentiableAverageLagging`` comes from
SimulEval (``simuleval.metrics.latency``, imported at mma_criterion.py:15-26; the reference pins no
version: docs/simuleval_instruction.md:6 installs master).  SimulEval is not installed here and
not vendored, so its published algorithm (Arivazhagan et al. 2019, eq. 20-21; SimulEval 1.0
``latency_metric`` wrapper semantics: delays [B,T], src_lens [B], ref_lens [B] optional,
target_padding_mask [B,T] optional; gamma = ref_len/src_len when ref_lens is given) is restated
below.  Parity of everything AROUND it is pinned: ``tests/golden/latency.npz`` is produced by
executing the reference's own ``compute_latency_loss`` source (ref_loader.load_mma_latency_loss)
with this restatement injected as ``LATENCY_METRICS``; the DAL restatement itself is "parity
unpinned" (no SimulEval copy to run) and says so here and in DESIGN.md.
"""
from types import SimpleNamespace
from typing import List, Optional

import torch
from torch import Tensor


def differentiable_average_lagging(delays: Tensor, src_lens: Tensor, ref_lens: Optional[Tensor] = None,
                                   target_padding_mask: Optional[Tensor] = None) -> Tensor:
    """SimulEval ``DifferentiableAverageLagging`` (restated): g'(1) = g(1),
    g'(i) = max(g(i), g'(i-1) + 1/gamma), DAL = 1/|Y| sum_i (g'(i) - (i-1)/gamma), gamma = |Y|/|X|.
    Returns [B, 1]."""
    assert delays.dim() == 2
    bsz, t_len = delays.shape
    src = src_lens.view(-1, 1).type_as(delays)
    if target_padding_mask is not None:
        tgt = (t_len - target_padding_mask.sum(dim=1)).view(-1, 1)
        delays = delays.masked_fill(target_padding_mask, 0)
    else:
        tgt = torch.ones_like(src) * t_len
    gamma = (ref_lens.view(-1, 1).type_as(delays) if ref_lens is not None else tgt) / src
    cols = [delays[:, 0]]
    for i in range(1, t_len):
        cols.append(torch.maximum(cols[-1] + (1 / gamma).view(-1), delays[:, i]))
    new_delays = torch.stack(cols, dim=1)
    dal = new_delays - torch.arange(t_len).unsqueeze(0).type_as(delays).expand_as(delays) / gamma
    if target_padding_mask is not None:
        dal = dal.masked_fill(target_padding_mask, 0)
    return dal.sum(dim=1, keepdim=True) / tgt


LATENCY_METRICS = {"differentiable_average_lagging": differentiable_average_lagging}


def criterion_stub(latency_avg_weight=0.1, latency_var_weight=0.1, latency_gather_method="weighted_average",
                   padding_idx=1, ms_per_frame_shift=10.0):
    """The attributes ``compute_latency_loss`` reads from ``self`` (mma_criterion.py:81-96)."""
    return SimpleNamespace(latency_avg_weight=latency_avg_weight, latency_var_weight=latency_var_weight,
                           latency_avg_type="differentiable_average_lagging",
                           latency_gather_method=latency_gather_method, padding_idx=padding_idx,
                           ms_per_frame_shift=ms_per_frame_shift)


def mma_latency_loss(alpha_list: List[Tensor], target: Tensor, src_lengths: Tensor,
                     encoder_padding_mask: Tensor, cfg) -> tuple:
    """Restatement of mma_criterion.py:138-207.  alpha_list: per layer [bsz, heads, T, S].
    Returns (latency_loss, expected_latency.sum(), expected_delays_var, expected_delays [N,T])."""
    num_layers = len(alpha_list)
    bsz, num_heads, tgt_len, src_len = alpha_list[0].shape
    alpha_all = torch.cat(alpha_list, dim=1).view(-1, tgt_len, src_len)
    steps = torch.arange(1, 1 + src_len).view(1, 1, -1).expand_as(alpha_all).type_as(alpha_all)
    expected_delays = torch.sum(steps * alpha_all, dim=-1)
    loss, latency, var = latency_from_delays(expected_delays, num_layers * num_heads, target, src_lengths,
                                             encoder_padding_mask, cfg)
    return loss, latency, var, expected_delays


def latency_from_delays(expected_delays: Tensor, heads_total: int, target: Tensor, src_lengths: Tensor,
                        encoder_padding_mask: Tensor, cfg) -> tuple:
    """mma_criterion.py:159-207 given the [bsz*layers*heads, T] expected delays."""
    bsz = target.shape[0]
    tgt_len = expected_delays.shape[1]
    target_padding_mask = target == cfg.padding_idx
    target_lengths = (~target_padding_mask).sum(1)
    assert not bool(encoder_padding_mask[:, 0].any()), "Only right padding is supported."
    encoder_lengths = (~encoder_padding_mask).sum(-1)

    def expand(t):
        return torch.repeat_interleave(t, heads_total, 0)

    expected_latency = differentiable_average_lagging(
        expected_delays, expand(encoder_lengths), expand(target_lengths),
        target_padding_mask=expand(target_padding_mask))
    expected_latency = expected_latency.view(bsz, -1)
    if cfg.latency_gather_method == "average":
        expected_latency = expected_delays.mean(dim=1)          # sic (:184)
    elif cfg.latency_gather_method == "weighted_average":
        weights = torch.softmax(expected_latency, dim=1)
        expected_latency = torch.sum(expected_latency * weights, dim=1)
    elif cfg.latency_gather_method == "max":
        expected_latency = expected_latency.max(dim=1)[0]
    else:
        raise NotImplementedError
    avg_loss = cfg.latency_avg_weight * expected_latency.clip(min=0).sum()
    expected_delays_var = expected_delays.view(bsz, -1, tgt_len).var(dim=1).mean(dim=1).sum()
    latency_loss = avg_loss + cfg.latency_var_weight * expected_delays_var
    expected_latency = expected_latency * (src_lengths / encoder_lengths * cfg.ms_per_frame_shift)
    return latency_loss, expected_latency.sum(), expected_delays_var
