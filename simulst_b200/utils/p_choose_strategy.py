"""Mirror of ``codebase/utils/p_choose_strategy.py`` of the reference."""
from typing import Dict, Optional

import torch
from torch import Tensor

from .. import ops


def waitk_p_choose(
    tgt_len: int,
    src_len: int,
    bsz: int,
    waitk_lagging: int,
    key_padding_mask: Optional[Tensor] = None,
    incremental_state: Optional[Dict[str, Dict[str, Optional[Tensor]]]] = None
):
    """p_choose_strategy.py:6-53 -- wait-k policy: target step i writes after reading source
    frame ``min(i + k - 1, eos)`` (no ``min`` when ``incremental_state["online"]``).

    Same signature, result (bool ``[bsz, 1, src_len]``, device of the mask -- CPU without one,
    :25) and failure mode as the reference: ``incremental_state`` is dereferenced
    unconditionally (:35), so calling it without one (wait-k *training*) raises AttributeError
    upstream and does so here -- SURVEY Appendix Q; not silently "fixed".  Because that makes the
    reference's final ``[:, -1:]`` slice (:50-51) unconditional too, only the LAST target row is
    ever returned, and only that row is generated here: one comparison of ``arange(src_len)``
    against a ``[bsz, 1]`` step index instead of a ``[bsz, tgt_len, src_len]`` one-hot.
    Integer index generation on ``bsz * src_len`` elements: not a kernel."""
    online = incremental_state.get("online", False)
    if key_padding_mask is None:
        last_frame = torch.full((bsz,), src_len - 1)
    else:
        last_frame = key_padding_mask.logical_not().long().sum(-1) - 1
    dev = last_frame.device
    step = torch.full((bsz, 1), tgt_len - 1 + waitk_lagging - 1, dtype=torch.long, device=dev)
    if not online:
        step = torch.minimum(step, last_frame.view(bsz, 1))
    return (torch.arange(src_len, device=dev).view(1, 1, src_len) == step.view(bsz, 1, 1))


def learnable_p_choose(
    energy,
    noise_mean: float = 0.0,
    noise_std: float = 1.0,
    training: bool = True
):
    """p_choose_strategy.py:56-76: sigmoid(energy + N(mean, std) noise in training).  The noise
    is drawn with torch's generator (same stream of random numbers as the reference); the add
    and the sigmoid are one kernel."""
    noise = None
    if training:
        noise = torch.randn_like(energy) * noise_std + noise_mean
    return ops.PChooseFunction.apply(energy, noise)
