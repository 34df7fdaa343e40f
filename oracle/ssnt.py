"""Oracle (CPU PyTorch) for the SSNT lattice loss -- SURVEY 8f rank 4.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Reference: ``codebase/criterion/ssnt_loss/ssnt_loss.py`` (git submodule ssnt_loss @ a5af91e)
  :14-26    log_exclusive_cumprod / exclusive_cumsum
  :45-151   ssnt_loss            (padded [N, T, S(, V)] layout; recurrence at :121-127)
  :154-271  ssnt_loss_mem        (targets concatenated: [T_flat, S(, V)])
and its own acceptance test ``ssnt_loss/test.py:19-80`` (an O(T*S^2) triple loop, restated here as
``ssnt_lattice_bruteforce``), which pins the oracle next to the golden vectors generated from the
unmodified reference (tests/golden/ssnt.npz).

The loss is Monotonic Attention's expected alignment in log space with the word-prediction
probability folded in:
    log_alpha[0]   = [0, -inf, -inf, ...]
    log_alpha[i+1] = clamp( logp_trans[i] + log_p[i] + lcp[i]
                            + logcumsumexp(log(1 + lambda) + log_alpha[i] - lcp[i]),  neg_inf, 0 )
    lcp[i]         = exclusive cumsum over the source axis of log(1 - p[i])
    loss[n]        = -log_alpha[n, target_len[n], source_len[n] - 1]
"""
from typing import Optional

import torch
from torch import Tensor

F32, F64 = torch.float32, torch.float64


def _excl_cumsum_last(x: Tensor) -> Tensor:
    """ssnt_loss.py:22-26 along the last axis: shift right by one, zero in front, cumsum."""
    return torch.cat([torch.zeros_like(x[..., :1]), x[..., :-1]], dim=-1).cumsum(-1)


def _emission_logs(emit_logits, emit_probs, dt):
    if emit_logits is not None:
        z = emit_logits.to(dt) if dt == F64 else emit_logits
        return torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
    assert emit_probs is not None, "emit_probs and emit_logits cannot both be None."
    p = emit_probs.to(dt) if dt == F64 else emit_probs
    return torch.log(p), torch.log1p(-p)


def ssnt_loss(log_probs: Tensor, targets: Tensor, source_lengths: Tensor, target_lengths: Tensor,
              emit_logits: Optional[Tensor] = None, emit_probs: Optional[Tensor] = None,
              neg_inf: float = -1e4, reduction: str = "none", fastemit_lambda: float = 0.0,
              compute_dtype: torch.dtype = F32):
    """ssnt_loss.py:45-151.  Returns (loss, lattice [N,T,S], log_p_choose [N,T,S])."""
    dt = compute_dtype
    log_p, log_1mp = _emission_logs(emit_logits, emit_probs, dt)
    n, t_len, s_len = log_p.shape
    log_p = log_p.to(dt)
    pad = torch.arange(s_len).view(1, s_len) >= source_lengths.view(n, 1)
    if bool(pad.any()):
        log_p = log_p.masked_fill(pad.unsqueeze(1), neg_inf)
    lcp = _excl_cumsum_last(log_1mp)
    trans = log_probs.gather(-1, targets.view(n, t_len, 1, 1).expand(-1, -1, s_len, -1)).squeeze(-1)
    if dt == F64:
        trans, lcp = trans.to(dt), lcp.to(dt)
    prefix = trans + log_p + lcp
    fe = torch.tensor([fastemit_lambda]).log1p().to(dt)
    rows = [torch.cat([log_p.new_zeros(n, 1), log_p.new_full((n, s_len - 1), neg_inf)], dim=1)]
    for i in range(t_len):
        nxt = prefix[:, i] + torch.logcumsumexp(fe + rows[-1] - lcp[:, i], dim=1)
        rows.append(nxt.clamp(min=neg_inf, max=0))
    log_alpha = torch.stack(rows, dim=1)
    lattice = log_alpha[:, 1:]
    at_src_end = log_alpha.gather(2, (source_lengths - 1).view(n, 1, 1).expand(-1, 1 + t_len, -1))
    ll = at_src_end.gather(1, target_lengths.view(n, 1, 1)).view(n)
    if reduction == "sum":
        ll = ll.sum()
    elif reduction == "mean":
        ll = ll.mean()
    return -ll, lattice, log_p


def ssnt_loss_mem(log_probs: Tensor, targets: Tensor, source_lengths: Tensor, target_lengths: Tensor,
                  emit_logits: Optional[Tensor] = None, emit_probs: Optional[Tensor] = None,
                  neg_inf: float = -1e4, reduction: str = "none", fastemit_lambda: float = 0.0,
                  compute_dtype: torch.dtype = F32):
    """ssnt_loss.py:154-271: the same recurrence on targets concatenated over the batch
    (log_probs [T_flat, S, V], emit [T_flat, S]); lattice rows: sample n owns rows
    off_out[n] .. off_out[n] + target_len[n] of a [T_flat + N, S] buffer, the first being alpha_0."""
    dt = compute_dtype
    log_p, log_1mp = _emission_logs(emit_logits, emit_probs, dt)
    n = source_lengths.shape[0]
    t_flat, s_len = log_p.shape
    log_p = log_p.to(dt)
    src_rep = torch.repeat_interleave(source_lengths, target_lengths, dim=0)
    pad = torch.arange(s_len).view(1, s_len) >= src_rep.view(-1, 1)
    if bool(pad.any()):
        log_p = log_p.masked_fill(pad, neg_inf)
    lcp = _excl_cumsum_last(log_1mp)
    off = torch.cumsum(target_lengths, 0) - target_lengths
    off_out = torch.cumsum(target_lengths + 1, 0) - (target_lengths + 1)
    trans = log_probs.gather(-1, targets.view(t_flat, 1, 1).expand(-1, s_len, -1)).squeeze(-1)
    if dt == F64:
        trans, lcp = trans.to(dt), lcp.to(dt)
    prefix = trans + log_p + lcp
    fe = torch.tensor([fastemit_lambda]).log1p().to(dt)
    log_alpha = log_p.new_zeros(t_flat + n, s_len)
    log_alpha[off_out, 1:] = neg_inf
    for i in range(int(target_lengths.max())):
        live = i < target_lengths
        src, dst = (off + i)[live], (off_out + i)[live]
        log_alpha[dst + 1] = (prefix[src] + torch.logcumsumexp(fe + log_alpha[dst] - lcp[src], dim=1)
                              ).clamp(min=neg_inf, max=0)
    lattice = log_alpha
    ll = log_alpha[off_out + target_lengths].gather(-1, (source_lengths - 1).view(n, 1)).view(n)
    if reduction == "sum":
        ll = ll.sum()
    elif reduction == "mean":
        ll = ll.mean()
    return -ll, lattice, log_p


def ssnt_lattice_bruteforce(log_probs: Tensor, targets: Tensor, source_lengths: Tensor,
                            target_lengths: Tensor, emit_probs: Tensor, neg_inf: float = -1e4):
    """The reference's acceptance checker (ssnt_loss/test.py:19-80), restated: explicit sum over
    the previous frame k <= i of   alpha[j-1, k] * p[j, i] * prod_{k <= m < i} (1 - p[j, m]),
    in fp64 log space.  O(T*S^2) per sample -- tiny cases only.  Returns (loss [N], lattice [N,T,S])."""
    lp = torch.log(emit_probs.double())
    l1 = torch.log1p(-emit_probs.double())
    n, t_len, s_len = lp.shape
    pad = torch.arange(s_len).view(1, s_len) >= source_lengths.view(n, 1)
    lp = lp.masked_fill(pad.unsqueeze(1), neg_inf)
    c = _excl_cumsum_last(l1)
    lat = lp.new_zeros(n, t_len, s_len)
    loss = lp.new_zeros(n)
    for b in range(n):
        for j in range(t_len):
            y = int(targets[b, j])
            for i in range(s_len):
                if j == 0:
                    inner = c[b, 0, i]               # alpha_0 sits on frame 0: skip frames 0..i-1
                else:
                    terms = [lat[b, j - 1, k] + c[b, j, i] - c[b, j, k] for k in range(i + 1)]
                    inner = torch.logsumexp(torch.stack(terms), 0)
                lat[b, j, i] = inner + lp[b, j, i] + log_probs[b, j, i, y].double()
        loss[b] = -lat[b, int(target_lengths[b]) - 1, int(source_lengths[b]) - 1]
    return loss, lat
