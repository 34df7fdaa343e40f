"""Mirror of ``codebase/models/torch_cif/cif.py`` of the reference: same signature, same
dict-of-lists return structure and dtypes; the integration runs in the sm_100a kernels."""
from typing import Optional

import torch
from torch import Tensor

from ... import ops

__all__ = ["cif_function", "prob_check"]


def prob_check(tensor, eps=1e-10, neg_inf=-1e8, logp=False):
    """cif.py:6-20 (host-syncing assertion, kept literal)."""
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


def cif_function(
    input: Tensor,
    alpha: Tensor,
    beta: float = 1.0,
    tail_thres: float = 0.5,
    padding_mask: Optional[Tensor] = None,
    target_lengths: Optional[Tensor] = None,
    eps: float = 1e-4,
):
    r"""A parallel implementation of continuous integrate-and-fire (CIF)
    https://arxiv.org/abs/1905.11235 -- interface of the reference's ``cif_function``
    (cif.py:23-196).

    Args:
        input (Tensor): (N, S, C) Input features to be integrated.
        alpha (Tensor): (N, S) Weights corresponding to each elements in the input (after sigmoid).
        beta (float): the threshold used for determine firing.
        tail_thres (float): the threshold for determine firing for tail handling.
        padding_mask (Tensor, optional): (N, S) binary mask of padded elements.
        target_lengths (Tensor, optional): (N,) desired target lengths (training mode).
        eps (float, optional): Epsilon to prevent underflow for divisions. Default: 1e-4

    Returns -> Dict[str, List[Optional[Tensor]]]: cif_out (N, T, C), cif_lengths (N,),
        alpha_sum (N,), delays (N, T), tail_weights (N,) (inference only).
    """
    B, S, C = input.size()
    assert tuple(alpha.size()) == (B, S), f"{alpha.size()} != {(B, S)}"
    if padding_mask is not None:
        padding_mask = padding_mask.bool()
    out, delays, alpha_sum, lengths, tail = ops.CIFFunction.apply(
        input, alpha, padding_mask, target_lengths, float(beta), float(tail_thres), float(eps))
    return {
        "cif_out": [out],
        "cif_lengths": [lengths],
        "alpha_sum": [alpha_sum],
        "delays": [delays],
        "tail_weights": [tail] if target_lengths is None else []
    }
