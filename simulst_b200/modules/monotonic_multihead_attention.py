"""Drop-in bodies for ``MonotonicAttention.monotonic_attention_process_train`` and
``monotonic_attention_process_infer`` of the reference
(codebase/modules/monotonic_multihead_attention.py:152-352).

The reference classes derive from fairseq's ``MultiheadAttention`` and are chosen through the
``--simul-attn-type`` registry; everything around these two methods (projections, energy bmm's,
``forward``, registry, state re-ordering) is unchanged.  ``B200MonotonicAttentionMixin`` is
placed in front of the reference class::

    class MonotonicAttention(B200MonotonicAttentionMixin, _ReferenceMonotonicAttention): ...

or applied to an existing class with ``patch_monotonic_attention(cls)``.  The mixin only uses
attributes the reference classes already have: ``p_choose``, ``energy_from_qk``, ``eps``,
``mass_preservation``, ``soft_attention``, ``chunk_size``, ``num_heads``,
``_get_monotonic_buffer`` / ``_set_monotonic_buffer``.
"""
from typing import Dict, Optional

import torch
from torch import Tensor

from .. import ops
from ..utils.monotonic_attention import mma_process_train, mma_process_train_with_delays


class B200MonotonicAttentionMixin:
    def monotonic_attention_process_train(
        self,
        query: Optional[Tensor],
        key: Optional[Tensor],
        key_padding_mask: Optional[Tensor] = None,
    ):
        """reference :301-352 -- p_choose, then ONE fused launch for expected alignment,
        mass preservation and expected soft attention (fp32 alpha / beta, as the reference's
        ``p_choose.float()`` path yields)."""
        assert query is not None
        assert key is not None

        # 1. compute stepwise probability
        p_choose = self.p_choose(query, key, key_padding_mask)

        # 2./3. expected alignment (+ mass preservation) (+ expected soft attention)
        if self.soft_attention:
            soft_energy = self.energy_from_qk(
                query,
                key,
                "soft",
                key_padding_mask=None,
            )
            alpha, beta, delays = self._alignment(p_choose, soft_energy, key_padding_mask, self.chunk_size)
        else:
            alpha, beta, delays = self._alignment(p_choose, None, key_padding_mask, None)
            soft_energy = alpha

        # [bsz*heads, tgt] expected delays of this layer (sum_j (j+1)*alpha) when the module has
        # `with_expected_delays = True`: what MMACriterion.compute_latency_loss
        # (mma_criterion.py:146-157) otherwise recomputes from alpha.  Kept on the module so a
        # criterion can use it instead of re-reading alpha; the reference's return tuple is unchanged.
        self.expected_delays = delays

        return p_choose, alpha, beta, soft_energy

    def _alignment(self, p_choose, soft_energy, key_padding_mask, chunk_size):
        if getattr(self, "with_expected_delays", False):
            return mma_process_train_with_delays(
                p_choose, soft_energy, key_padding_mask, eps=self.eps,
                mass_preservation=self.mass_preservation, chunk_size=chunk_size)
        alpha, beta = mma_process_train(
            p_choose, soft_energy, key_padding_mask, eps=self.eps,
            mass_preservation=self.mass_preservation, chunk_size=chunk_size)
        return alpha, beta, None

    def monotonic_attention_process_infer(
        self,
        query: Optional[Tensor],
        key: Optional[Tensor],
        key_padding_mask: Optional[Tensor] = None,
        incremental_state: Optional[Dict[str, Dict[str, Optional[Tensor]]]] = None,
    ):
        """reference :152-299 -- one kernel launch per layer and step, no host reads (the
        reference's two ``assert ....max()`` syncs are structural invariants here)."""
        assert query is not None
        assert key is not None

        tgt_len, bsz, _ = query.size()
        src_len = key.size(0)
        assert tgt_len == 1
        bsz_head = bsz * self.num_heads

        # 1. compute stepwise probability
        p_choose = self.p_choose(
            query, key, key_padding_mask, incremental_state
        ).squeeze(1)

        src_lengths = None
        if key_padding_mask is not None:
            assert key_padding_mask.size() == (bsz_head, src_len), (
                f"{key_padding_mask.size()} != {bsz_head, src_len}")
            src_lengths = (~key_padding_mask).sum(1)

        # 2. head positions carried between steps
        monotonic_cache = self._get_monotonic_buffer(incremental_state)
        head_step = monotonic_cache.get('head_step', None)
        if head_step is None:
            head_step = p_choose.new_zeros(bsz_head).long()
        head_step = head_step.reshape(bsz_head).contiguous().clone()

        soft_energy = None
        if self.soft_attention:
            soft_energy = self.energy_from_qk(
                query,
                key,
                "soft",
                key_padding_mask=key_padding_mask,
            ).squeeze(1)

        head_read, alpha, beta = ops.mma_step(
            p_choose, head_step, soft_energy, src_lengths, self.mass_preservation)

        monotonic_cache["head_step"] = head_step.view(bsz, self.num_heads)  # for reorder to work.
        # Whether a head is looking for new input
        monotonic_cache["head_read"] = head_read.view(bsz, self.num_heads)
        self._set_monotonic_buffer(incremental_state, monotonic_cache)

        if self.soft_attention:
            beta = beta.view(bsz_head, tgt_len, src_len)
        else:
            # If it's hard attention just select the last state
            beta = alpha.view(bsz_head, tgt_len, src_len)

        return p_choose, alpha, beta


def patch_monotonic_attention(cls):
    """Replace the two method bodies of an existing reference attention class in place (and add
    the helper they share).  Everything else of the class -- projections, energy bmm's,
    ``forward``, incremental-state plumbing -- stays the reference's."""
    cls.monotonic_attention_process_train = B200MonotonicAttentionMixin.monotonic_attention_process_train
    cls.monotonic_attention_process_infer = B200MonotonicAttentionMixin.monotonic_attention_process_infer
    cls._alignment = B200MonotonicAttentionMixin._alignment
    if not hasattr(cls, "expected_delays"):
        cls.expected_delays = None
    return cls
