"""Drop-in training path for the reference's ``*_fixed_pre_decision`` attention classes
(codebase/modules/fixed_pre_decision.py:175-190, the configuration ``exp/2-mma.sh:56-57`` trains:
``infinite_lookback_fixed_pre_decision`` with ratio 8).

The reference computes p_choose on the POOLED keys (``[N,T,ceil(S/ratio)]``, :107-137), blows it
up to ``[N,T,S]`` with a transposed convolution (``insert_zeros``, :85-95), patches the last
column (:139-159) and then runs the dense alignment functions over a tensor that is (ratio-1)/ratio
zeros.  Here the pooled tensor goes straight to ``simulst_mma_train_fwd_pooled``: alpha is zero off
the pooled grid, so the recurrence runs on ``[N,T,ceil(S/ratio)]`` only, streaming row kernels produce
the soft attention and the dense outputs (csrc/mma_sparse.cu), and the backward returns the gradient of
the pooled tensor.

``B200FixedStrideMixin`` goes in front of a class produced by the reference's
``fixed_pooling_monotonic_attention`` decorator (or apply ``patch_fixed_pre_decision(cls)``); it
uses only attributes those classes have: ``pooling_layer``, ``pre_decision_ratio``,
``pre_decision_pad_threshold``, ``p_choose_from_qk``, ``energy_from_qk``.  Inference
(``monotonic_attention_process_infer``) keeps the reference's ``p_choose`` -- one decoding step
expands a single ``[N,1,S]`` row -- followed by the ``simulst_mma_step`` body of
``B200MonotonicAttentionMixin``.
"""
from typing import Optional

from torch import Tensor

from .. import ops
from .monotonic_multihead_attention import B200MonotonicAttentionMixin


class B200FixedStrideMixin(B200MonotonicAttentionMixin):
    # The dense p_choose is part of the reference's return value (forward() puts it into the
    # attention dict, :417-421).  Nothing in the reference reads it afterwards; a caller that
    # does not either sets this to False and the [N,T,S] tensor is never written.
    return_dense_p_choose = True
    # Likewise the dense [N,T,S] alpha: with `with_expected_delays = True` the latency loss reads
    # the [N,T] delays instead (mma_criterion.py:146-157) and soft attention reaches the output
    # through beta.  False (with return_dense_p_choose = False, soft attention, a shape the
    # pooled-grid kernels take): alpha is returned as None and never written.
    return_dense_alpha = True
    # forward() asserts "Only right padding is supported." (monotonic_multihead_attention.py:378-381);
    # the kernels take that as a promise and VERIFY it per row (a violation poisons the row's
    # outputs with NaN and sets SIMULST_ST_NOT_RIGHT_PADDED).  False: masked calls expand the row
    # and run the arbitrary-mask path.
    assume_right_padding = True

    def p_choose_pooled(self, query: Optional[Tensor], key: Optional[Tensor],
                        key_padding_mask: Optional[Tensor] = None):
        """reference :96-137 (training branch, incremental_state None): keys and padding mask
        pooled with the wrapper's own pooling layer, p_choose from the pooled keys."""
        assert key is not None
        assert query is not None
        key_pool = self.pooling_layer(key.transpose(0, 2)).transpose(0, 2)
        if key_padding_mask is not None:
            key_padding_mask_pool = (
                self.pooling_layer(key_padding_mask.unsqueeze(0).float())
                .squeeze(0)
                .gt(self.pre_decision_pad_threshold)
            )
            # Make sure at least one element is not pad
            key_padding_mask_pool[:, 0] = 0
        else:
            key_padding_mask_pool = None
        return self.p_choose_from_qk(query, key_pool, key_padding_mask_pool, incremental_state=None)

    def monotonic_attention_process_train(
        self,
        query: Optional[Tensor],
        key: Optional[Tensor],
        key_padding_mask: Optional[Tensor] = None,
    ):
        """reference monotonic_multihead_attention.py:301-352 with p_choose() of
        fixed_pre_decision.py:96-170 folded in: the pooled p_choose goes to the kernels as it is."""
        assert query is not None
        assert key is not None
        src_len = key.size(0)
        p_pooled = self.p_choose_pooled(query, key, key_padding_mask)
        assert p_pooled.size(-1) * self.pre_decision_ratio >= src_len
        soft_energy = None
        if self.soft_attention:
            soft_energy = self.energy_from_qk(query, key, "soft", key_padding_mask=None)
        p_choose, alpha, beta, delays = ops.mma_train_pooled(
            p_pooled, src_len, self.pre_decision_ratio, soft_energy, key_padding_mask, eps=self.eps,
            mass_preservation=self.mass_preservation,
            chunk_size=self.chunk_size if self.soft_attention else None,
            with_delays=bool(getattr(self, "with_expected_delays", False)),
            want_dense=self.return_dense_p_choose, right_padding=self.assume_right_padding,
            want_alpha=self.return_dense_alpha)
        self.expected_delays = delays
        if not self.soft_attention:
            soft_energy = alpha
        return p_choose, alpha, beta, soft_energy


def patch_fixed_pre_decision(cls):
    """Replace the training/inference method bodies of an existing ``*_fixed_pre_decision`` class
    in place; everything else (pooling layer, projections, ``p_choose`` for inference,
    ``forward``) stays the reference's."""
    cls.p_choose_pooled = B200FixedStrideMixin.p_choose_pooled
    cls.monotonic_attention_process_train = B200FixedStrideMixin.monotonic_attention_process_train
    cls.monotonic_attention_process_infer = B200MonotonicAttentionMixin.monotonic_attention_process_infer
    cls._alignment = B200MonotonicAttentionMixin._alignment
    for name in ("return_dense_p_choose", "return_dense_alpha", "assume_right_padding"):
        if not hasattr(cls, name):
            setattr(cls, name, getattr(B200FixedStrideMixin, name))
    if not hasattr(cls, "expected_delays"):
        cls.expected_delays = None
    return cls
