"""Oracle (CPU PyTorch) for Continuous Integrate-and-Fire.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Reference:
  /root/reference/codebase/models/torch_cif/cif.py:23-196      cif_function
  /root/reference/codebase/models/torch_cif/test.py:24-91      sequential checker (the
      reference's own ground truth for its property test) -> ``cif_sequential`` here
  /root/reference/codebase/models/cif_transformer.py:143-186   CIFLayer.forward body
  /root/reference/codebase/models/cif_transformer.py:188-261   CIFLayer.infer body

``cif_function`` below is the parallel formulation (cumsum -> floor -> three families of
accumulating scatters), differentiable through autograd exactly like the reference:
indices are computed without gradient, weights carry gradient.
"""
from typing import Dict, List, Optional

import torch
from torch import Tensor

F32 = torch.float32


def _accumulate(buf: Tensor, rows: Tensor, src: Tensor) -> Tensor:
    """buf[b, rows[b, s]] += src[b, s]   (buf [B, T1, ...], rows [B, S] int64)."""
    b, t1 = buf.shape[:2]
    flat_rows = (rows + torch.arange(b).unsqueeze(1) * t1).reshape(-1)
    flat_buf = buf.reshape(b * t1, *buf.shape[2:])
    flat_src = src.reshape(-1, *src.shape[2:])
    return flat_buf.index_add(0, flat_rows, flat_src).reshape(buf.shape)


def cif_function(
    input: Tensor,
    alpha: Tensor,
    beta: float = 1.0,
    tail_thres: float = 0.5,
    padding_mask: Optional[Tensor] = None,
    target_lengths: Optional[Tensor] = None,
    eps: float = 1e-4,
    compute_dtype: torch.dtype = F32,
) -> Dict[str, List[Tensor]]:
    """cif.py:23-196.  Returns the same dict-of-lists.  ``compute_dtype=float64``
    evaluates weights/indices in double (inputs are up-cast)."""
    b_sz, s_len, c_dim = input.shape
    assert tuple(alpha.shape) == (b_sz, s_len), f"{alpha.shape} != {(b_sz, s_len)}"
    assert not bool(torch.isnan(alpha).any()), "Nan in a probability tensor."
    assert bool((alpha <= 1 + 1e-10).all()) and bool((alpha >= -1e-10).all()), (
        "Incorrect values in a probability tensor, 0.0 <= tensor <= 1.0")

    alpha_dtype = alpha.dtype
    x = input if compute_dtype == F32 else input.to(compute_dtype)
    a = alpha.to(compute_dtype)
    if padding_mask is not None:
        a = a.masked_fill(padding_mask.bool(), 0)

    training = target_lengths is not None
    if training:                                                # cif.py:66-72
        lengths = target_lengths.long()
        wanted = beta * target_lengths.type_as(x) + eps
        alpha_sum = a.sum(1)
        a = a * (wanted / alpha_sum).unsqueeze(1)
    else:                                                       # cif.py:73-76
        alpha_sum = a.sum(1)
        lengths = (alpha_sum / beta).floor().long()
    t_max = int(lengths.max())

    csum = a.cumsum(-1)                                         # cif.py:79
    with torch.no_grad():                                       # cif.py:80-88
        right = (csum / beta).floor().long().clamp(max=t_max)
        left = torch.cat([torch.zeros_like(right[:, :1]), right[:, :-1]], dim=1)
        fires = right - left
        extra = (fires - 1).clamp(min=0)

    out = x.new_zeros(b_sz, t_max + 1, c_dim)
    delay = x.new_zeros(b_sz, t_max + 1)
    pos = torch.arange(1, 1 + s_len).unsqueeze(0).type_as(x)

    # weight falling right of the last threshold crossed by frame s   (cif.py:96-112)
    w_right = torch.where(fires > 0, csum - right.type_as(a) * beta, a.new_zeros(1)).type_as(x)
    out = _accumulate(out, right, w_right.unsqueeze(-1) * x)
    delay = _accumulate(delay, right, w_right * pos / beta)

    # weight completing the segment that was open when frame s arrived  (cif.py:114-128)
    w_left = (a - w_right - extra.type_as(a) * beta).type_as(x)
    out = _accumulate(out, left, w_left.unsqueeze(-1) * x)
    delay = _accumulate(delay, left, w_left * pos / beta)

    # frames heavy enough to fill whole segments on their own           (cif.py:131-149)
    n_extra = int(extra.max()) if extra.numel() else 0
    tgt = left
    for k in range(1, n_extra + 1):
        tgt = (tgt + 1).clamp(max=t_max)
        live = (extra >= k)
        out = _accumulate(out, tgt, (x * beta) * live.unsqueeze(2))
        delay = _accumulate(delay, tgt, pos * live)

    if training:                                                # cif.py:152-155
        out = out[:, :t_max]
        delay = delay[:, :t_max]
        tail: List[Tensor] = []
    else:                                                       # cif.py:156-188
        zero = w_right.new_zeros(1)
        tail_w = torch.where(right == lengths.unsqueeze(1), w_right, zero).sum(-1)
        tail_w = tail_w + torch.where(left == lengths.unsqueeze(1), w_left, zero).sum(-1)
        grow = tail_w >= tail_thres
        if bool(grow.any()):
            factor = (beta / tail_w.masked_fill(~grow, beta)).detach()
            scale = torch.ones_like(out).scatter(
                1, lengths.view(b_sz, 1, 1).expand(-1, -1, c_dim),
                factor.view(b_sz, 1, 1).expand(-1, -1, c_dim)).detach()
            out = out * scale
            lengths = lengths + grow.long()
            t_max = int(lengths.max())
        out = out[:, :t_max]
        delay = delay[:, :t_max]
        dead = torch.arange(t_max).unsqueeze(0) >= lengths.unsqueeze(1)
        out = out.masked_fill(dead.unsqueeze(-1), 0)
        tail = [tail_w]

    return {
        "cif_out": [out],
        "cif_lengths": [lengths],
        "alpha_sum": [alpha_sum.to(alpha_dtype) if compute_dtype == F32 else alpha_sum],
        "delays": [delay],
        "tail_weights": tail,
    }


def cif_sequential(
    input: Tensor,
    alpha: Tensor,
    beta: float = 1.0,
    tail_thres: float = 0.5,
    padding_mask: Optional[Tensor] = None,
    target_lengths: Optional[Tensor] = None,
    eps: float = 1e-4,
):
    """Frame-by-frame integrate-and-fire (the ground truth of the reference's own
    property test, torch_cif/test.py:24-91), in double precision Python scalars.

    Returns (out [B, T(+1), C], delay [B, T(+1)]) with the reference checker's tail
    conventions: the slot holding the trailing partial segment is rescaled by
    beta/weight when weight >= tail_thres, otherwise zeroed together with everything
    after it; the extra slot T is dropped in training mode or when it is all-zero."""
    b_sz, s_len, c_dim = input.shape
    a = alpha.double().clone()
    x = input.double()
    if padding_mask is not None:
        a = a.masked_fill(padding_mask, 0)
        src_lengths = (~padding_mask).sum(-1)
    else:
        src_lengths = torch.full((b_sz,), s_len)
    if target_lengths is not None:
        lengths = target_lengths.long()
        a = a * ((beta * target_lengths.double() + eps) / a.sum(1)).unsqueeze(1)
    else:
        a32 = alpha.float()
        if padding_mask is not None:
            a32 = a32.masked_fill(padding_mask, 0)
        lengths = (a32.sum(1) / beta).floor().long()
    t_max = int(lengths.max())

    out = x.new_zeros(b_sz, t_max + 1, c_dim)
    delay = x.new_zeros(b_sz, t_max + 1)
    for b in range(b_sz):
        held = 0.0          # weight integrated into the currently open segment
        t = 0               # index of the open segment
        for s in range(int(src_lengths[b])):
            w = float(a[b, s])
            while held + w >= beta:
                part = beta - held
                out[b, t] += part * x[b, s]
                delay[b, t] += part * (s + 1) / beta
                w -= part
                held = 0.0
                t += 1
            held += w
            out[b, t] += w * x[b, s]
            delay[b, t] += w * (s + 1) / beta
        if held >= tail_thres:
            out[b, t] *= beta / held
        else:
            out[b, t:] = 0
    if target_lengths is not None or bool(out[:, t_max].eq(0).all()):
        out = out[:, :t_max]
        delay = delay[:, :t_max]
    return out, delay


# --------------------------------------------------------------------------- CIFLayer bodies
def cif_layer_forward(
    x: Tensor,                   # [S, B, C] encoder states
    alpha: Tensor,               # [B, S] integration weights AFTER sigmoid (alpha_proj is out of scope)
    beta: float,
    encoder_padding_mask: Optional[Tensor] = None,
    target_lengths: Optional[Tensor] = None,
) -> Dict[str, List[Tensor]]:
    """cif_transformer.py:157-186 from the point where alpha is a [B, S] probability."""
    x = x.transpose(1, 0)
    if encoder_padding_mask is not None:
        x = x.masked_fill(encoder_padding_mask.unsqueeze(2), 0)
        alpha = alpha.masked_fill(encoder_padding_mask, 0)
    res = cif_function(x, alpha, beta=beta, tail_thres=beta / 2,
                       target_lengths=target_lengths)
    res["cif_out"] = [res["cif_out"][0].transpose(0, 1)]
    res["alpha"] = [alpha]
    return res


def cif_layer_infer(
    x: Tensor,                   # [chunk, 1, C]
    alpha: Tensor,               # [1, chunk] after sigmoid
    state: Dict[str, Optional[Tensor]],    # carries prev_weight [1,1] / prev_feat [1,1,C]
    beta: float,
    finish: bool = False,
) -> Dict[str, List[Tensor]]:
    """cif_transformer.py:198-261 from the point where alpha is a probability.
    Mutates ``state`` like the reference mutates its incremental-state dict."""
    chunk, bsz, _ = x.shape
    if bsz > 1:
        raise NotImplementedError("batched infer not supported for now.")
    x = x.transpose(1, 0)
    if state.get("prev_weight") is not None and state["prev_weight"].numel() > 0:
        alpha = torch.cat((state["prev_weight"], alpha), dim=1)
        x = torch.cat((state["prev_feat"], x), dim=1)
    res = cif_function(x, alpha, beta=beta, tail_thres=(beta / 2) if finish else 0)
    feats = res["cif_out"][0]
    n_fired = res["cif_lengths"][0]
    tail_w = res["tail_weights"][0]
    if not finish:
        state["prev_feat"] = feats[:, int(n_fired) - 1:, :] / beta
        state["prev_weight"] = tail_w.view(bsz, 1)
    else:
        state["prev_feat"] = None
        state["prev_weight"] = None
    n_out = n_fired if finish else n_fired - 1
    res["cif_out"] = [feats.narrow(1, 0, int(n_out)).transpose(0, 1)]
    res["cif_lengths"] = [n_out]
    res["alpha"] = [alpha]
    return res
