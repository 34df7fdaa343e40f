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
    """p_choose_strategy.py:6-53: one-hot diagonal j == min(i + k - 1, eos).  Integer index
    generation only (no kernel needed).  Like the reference it dereferences
    ``incremental_state`` unconditionally (:35) -- wait-k *training* without an incremental
    state is broken upstream and is left so."""
    if key_padding_mask is not None:
        key_eos = (~key_padding_mask).long().sum(-1) - 1
    else:
        key_eos = torch.full((bsz,), src_len - 1)
    monotonic_step = (
        torch.arange(tgt_len, device=key_eos.device)
        .add(waitk_lagging - 1)
        .unsqueeze(0)
        .expand(bsz, -1)
        .clone()
    )
    online = incremental_state.get("online", False)
    if not online:
        monotonic_step = monotonic_step.clip(
            max=key_eos.unsqueeze(1).expand(-1, tgt_len)
        )
    p_choose = (
        torch.arange(src_len, device=key_eos.device)
        .unsqueeze(0)
        .unsqueeze(1)
        .expand(bsz, tgt_len, -1)
    ) == monotonic_step.unsqueeze(2)
    if incremental_state is not None:
        p_choose = p_choose[:, -1:]
    return p_choose


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
