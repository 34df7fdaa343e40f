"""Mirror of ``codebase/utils/functions.py`` of the reference: same names, arguments and error
behaviour, running on the sm_100a kernels (CUDA tensors only, no fallback)."""
import torch

from .. import _lib, ops


def prob_check(tensor, eps=1e-10):
    """functions.py:9-17.  Pure check with a host read: kept literal (the fused operators
    record the same conditions in the device status word instead, see simulst_b200.check_status)."""
    _lib.require_cuda(tensor)
    assert not torch.isnan(tensor).any(), (
        "Nan in a probability tensor."
    )
    assert tensor.le(1.0 + eps).all() and tensor.ge(0.0 - eps).all(), (
        "Incorrect values in a probability tensor"
        ", 0.0 <= tensor <= 1.0"
    )


def _last_dim_op(tensor, dim, fn):
    nd = tensor.dim()
    dim = dim % nd
    if dim == nd - 1:
        return fn(tensor)
    moved = tensor.transpose(dim, nd - 1).contiguous()
    return fn(moved).transpose(dim, nd - 1)


def _checked_cumprod(tensor, dim, eps, inclusive):
    """The reference reads ``(tensor + eps < 0).any().item()`` on every call (functions.py:57):
    same single host read here, on a status word of this call's own, so bits left by unrelated
    operators on the shared per-device word are neither consumed nor reported here."""
    word = torch.zeros(1, dtype=torch.int32, device=tensor.device)
    out = _last_dim_op(tensor, dim, lambda x: ops.exclusive_cumprod_lastdim(x, eps, inclusive=inclusive,
                                                                            status=word))
    bits = int(word.item())
    if bits:
        _lib.raise_for_status(bits)
    return out


def safe_cumprod(tensor, dim: int, eps: float = 1e-10):
    """functions.py:48-66: exp(cumsum(log(tensor + eps))) along `dim`; RuntimeError on
    tensor + eps < 0.  Differentiable, like the reference's composition."""
    _lib.require_cuda(tensor)
    return _checked_cumprod(tensor, dim, eps, True)


def exclusive_cumprod(tensor, dim: int, eps: float = 1e-10):
    """functions.py:20-45: [1, x1, x1x2, ...] (first element is exp(log(1+eps)))."""
    if dim not in (0, 1, 2):
        raise RuntimeError(
            "Cumprod on dimension 3 and more is not implemented"
        )
    _lib.require_cuda(tensor)
    return _checked_cumprod(tensor, dim, eps, False)


def moving_sum(x, start_idx: int, end_idx: int):
    """functions.py:69-125 over the last axis of a [N, T, S] tensor."""
    assert start_idx > 0 and end_idx > 0
    return ops.moving_sum(x, start_idx, end_idx)
