"""Oracle (CPU PyTorch) for Monotonic Multihead Attention's expected alignment.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function is written so
that with ``compute_dtype=torch.float32`` it performs the same sequence of torch
primitives as the reference (hence bit-comparable on CPU) and with
``torch.float64`` it evaluates the same formulas in double precision.

Reference (all under /root/reference/codebase/utils/):
  functions.py:9-17      prob_check
  functions.py:20-45     exclusive_cumprod
  functions.py:48-66     safe_cumprod
  functions.py:69-125    moving_sum
  monotonic_attention.py:12-76    expected_alignment_from_p_choose
  monotonic_attention.py:79-152   expected_soft_attention
  monotonic_attention.py:155-197  mass_preservation
  p_choose_strategy.py:6-53       waitk_p_choose
  p_choose_strategy.py:56-76      learnable_p_choose
"""
from typing import Optional

import torch
from torch import Tensor

F32 = torch.float32
F64 = torch.float64


# --------------------------------------------------------------------------- checks
def prob_check(t: Tensor, eps: float = 1e-10) -> None:
    """functions.py:9-17 -- NaN / range assertion (AssertionError on failure)."""
    assert not bool(torch.isnan(t).any()), "Nan in a probability tensor."
    ok = bool((t <= 1.0 + eps).all()) and bool((t >= 0.0 - eps).all())
    assert ok, "Incorrect values in a probability tensor, 0.0 <= tensor <= 1.0"


# --------------------------------------------------------------------------- scans
def safe_cumprod(t: Tensor, dim: int, eps: float = 1e-10) -> Tensor:
    """functions.py:48-66 -- exp(cumsum(log(t + eps))); RuntimeError if t + eps < 0."""
    if bool((t + eps < 0).any()):
        raise RuntimeError(
            "Safe cumprod can only take non-negative tensors as input."
        )
    return torch.exp(torch.cumsum(torch.log(t + eps), dim))


def exclusive_cumprod(t: Tensor, dim: int, eps: float = 1e-10) -> Tensor:
    """functions.py:20-45 -- [1, x1, x1x2, ...] built by prepending a ones slice.

    Because eps is added to the prepended one as well, element 0 is exp(log(1+eps)),
    i.e. 1.00000095 for eps=1e-6 in fp32, not 1.
    """
    if dim not in (0, 1, 2):
        raise RuntimeError("Cumprod on dimension 3 and more is not implemented")
    lead_shape = list(t.shape)
    lead_shape[dim] = 1
    padded = torch.cat([torch.ones(lead_shape, dtype=t.dtype, device=t.device), t], dim=dim)
    full = safe_cumprod(padded, dim=dim, eps=eps)
    return full.narrow(dim, 0, t.shape[dim])


def moving_sum(x: Tensor, start_idx: int, end_idx: int) -> Tensor:
    """functions.py:69-125 -- out[n] = sum_{m=n-start+1}^{n+end-1} x[m] along the last axis.

    The reference uses conv1d with a ones kernel; so does the oracle, to keep the
    same summation order.  x is [N, T, S].
    """
    assert start_idx > 0 and end_idx > 0
    n, t, s = x.shape
    width = start_idx + end_idx - 1
    flat = x.reshape(-1, s).unsqueeze(1)
    ones = torch.ones(1, 1, width, dtype=x.dtype, device=x.device)
    full = torch.nn.functional.conv1d(flat, ones, padding=width).squeeze(1)
    out = full[:, end_idx:-start_idx]
    assert out.shape[1] == s
    return out.reshape(n, t, s)


# --------------------------------------------------------------------------- a4
def expected_alignment_from_p_choose(
    p_choose: Tensor,
    padding_mask: Optional[Tensor] = None,
    eps: float = 1e-6,
    compute_dtype: torch.dtype = F32,
) -> Tensor:
    """monotonic_attention.py:12-76.

    alpha_0 = onehot(0); alpha_i = clamp(p_i * cp_i * cumsum(alpha_{i-1} / clamp(cp_i, eps, 1)), 0, 1)
    with cp_i = exclusive_cumprod(1 - p_i).  The unclamped cp is used in the
    prefix, the clamped one in the divisor (lines 46-47, 53, 59-64).
    Output is cast back to p_choose.dtype (line 72) unless computing in fp64.
    """
    prob_check(p_choose)
    n, t_len, s_len = p_choose.shape
    in_dtype = p_choose.dtype
    p = p_choose.to(compute_dtype)
    if padding_mask is not None:
        p = p.masked_fill(padding_mask.unsqueeze(1), 0.0)

    cp = exclusive_cumprod(1 - p, dim=2, eps=eps)
    cp_div = torch.clamp(cp, eps, 1.0)
    prefix = p * cp

    prev = p.new_zeros(n, s_len)
    prev[:, 0] = 1.0
    rows = []
    for i in range(t_len):
        prev = (prefix[:, i] * torch.cumsum(prev / cp_div[:, i], dim=1)).clamp(0, 1.0)
        rows.append(prev)
    alpha = torch.stack(rows, dim=1)
    if compute_dtype == F32:
        alpha = alpha.to(in_dtype)
    prob_check(alpha)
    return alpha


# --------------------------------------------------------------------------- a5
def mass_preservation(
    alpha: Tensor,
    padding_mask: Optional[Tensor] = None,
    left_padding: bool = False,
) -> Tensor:
    """monotonic_attention.py:155-197.  Out of place here (the reference mutates
    its argument when there is no mask; the mirror in simulst_b200 keeps that)."""
    prob_check(alpha)
    if padding_mask is not None:
        if not left_padding:
            assert not bool(padding_mask[:, 0].any()), (
                "Find padding on the beginning of the sequence."
            )
        alpha = alpha.masked_fill(padding_mask.unsqueeze(1), 0.0)

    if left_padding or padding_mask is None:
        residual = 1 - alpha[:, :, :-1].sum(dim=-1).clamp(0, 1)
        alpha = torch.cat([alpha[:, :, :-1], residual.unsqueeze(-1)], dim=-1)
    else:
        t_len = alpha.shape[1]
        residual = 1 - alpha.sum(dim=-1, keepdim=True).clamp(0, 1)
        last = (~padding_mask).sum(dim=1, keepdim=True) - 1
        alpha = alpha.scatter_add(2, last.expand(-1, t_len).unsqueeze(2), residual)
        prob_check(alpha)
    return alpha


# --------------------------------------------------------------------------- a7
def expected_soft_attention(
    alpha: Tensor,
    soft_energy: Tensor,
    padding_mask: Optional[Tensor] = None,
    chunk_size: Optional[int] = None,
    eps: float = 1e-10,
    compute_dtype: torch.dtype = F32,
) -> Tensor:
    """monotonic_attention.py:79-152 (infinite lookback: 128-137; chunkwise: 117-127)."""
    if padding_mask is not None:
        alpha = alpha.masked_fill(padding_mask.unsqueeze(1), 0.0)
        fill = -1e4 if soft_energy.dtype == torch.float16 else -1e8
        soft_energy = soft_energy.masked_fill(padding_mask.unsqueeze(1), fill)
    prob_check(alpha)
    out_dtype = alpha.dtype

    a = alpha.to(compute_dtype)
    e = soft_energy.to(compute_dtype)
    e = e - e.max(dim=2, keepdim=True)[0]
    ex = torch.exp(e) + eps

    if chunk_size is not None:
        inner = a / (eps + moving_sum(ex, chunk_size, 1))
        beta = ex * moving_sum(inner, 1, chunk_size)
    else:
        inner = a / (eps + torch.cumsum(ex, dim=2))
        beta = ex * torch.cumsum(inner.flip(dims=[2]), dim=2).flip(dims=[2])

    if padding_mask is not None:
        beta = beta.masked_fill(padding_mask.unsqueeze(1).to(torch.bool), 0.0)
    if compute_dtype == F32:
        beta = beta.to(out_dtype)
    beta = beta.clamp(0, 1)
    prob_check(beta)
    return beta


# --------------------------------------------------------------------------- a1 / a11
def learnable_p_choose(energy: Tensor, noise: Optional[Tensor] = None) -> Tensor:
    """p_choose_strategy.py:56-76 with the Gaussian noise supplied by the caller
    (already scaled: randn * std + mean), so oracle and kernel see the same draw."""
    return torch.sigmoid(energy if noise is None else energy + noise)


def waitk_p_choose(tgt_len: int, src_len: int, bsz: int, waitk_lagging: int,
                   key_padding_mask: Optional[Tensor] = None,
                   online: bool = False, last_only: bool = False) -> Tensor:
    """p_choose_strategy.py:6-53: one-hot at j == min(i + k - 1, eos) (no min when online)."""
    if key_padding_mask is not None:
        eos = (~key_padding_mask).long().sum(-1) - 1
    else:
        eos = torch.full((bsz,), src_len - 1)
    step = (torch.arange(tgt_len) + (waitk_lagging - 1)).unsqueeze(0).expand(bsz, -1).clone()
    if not online:
        step = torch.minimum(step, eos.unsqueeze(1).expand(-1, tgt_len))
    p = torch.arange(src_len).view(1, 1, -1).expand(bsz, tgt_len, -1) == step.unsqueeze(2)
    return p[:, -1:] if last_only else p


# --------------------------------------------------------------------------- a8
def mma_process_train(
    p_choose: Tensor,
    soft_energy: Optional[Tensor],
    padding_mask: Optional[Tensor] = None,
    eps: float = 1e-6,
    mass_preserve: bool = True,
    chunk_size: Optional[int] = None,
    compute_dtype: torch.dtype = F32,
):
    """Body of MonotonicAttention.monotonic_attention_process_train
    (modules/monotonic_multihead_attention.py:301-352) after the two energy bmm's:
    alpha = expected_alignment(p.float()) -> mass_preservation -> expected_soft_attention.
    Returns (alpha, beta); beta is alpha for the hard-aligned variant (soft_energy None)."""
    p32 = p_choose.float() if compute_dtype == F32 else p_choose.to(compute_dtype)
    alpha = expected_alignment_from_p_choose(p32, padding_mask, eps=eps,
                                             compute_dtype=compute_dtype)
    if mass_preserve:
        alpha = mass_preservation(alpha, padding_mask)
    if soft_energy is None:
        return alpha, alpha
    beta = expected_soft_attention(alpha, soft_energy, padding_mask=padding_mask,
                                   chunk_size=chunk_size, eps=eps,
                                   compute_dtype=compute_dtype)
    return alpha, beta


# --------------------------------------------------------------------------- 8f #2 (fixed pre-decision)
def insert_zeros(x: Tensor, stride: int) -> Tensor:
    """FixedStrideMonotonicAttention.insert_zeros (modules/fixed_pre_decision.py:85-95): a
    transposed 1-D convolution with the kernel [0, ..., 0, 1] of width `stride` and that stride,
    i.e. x[..., k] lands on column (k+1)*stride - 1 of a row of length Sp*stride."""
    n, t, sp = x.shape
    weight = torch.nn.functional.pad(torch.ones(1, 1, 1).to(x), (stride - 1, 0))
    up = torch.nn.functional.conv_transpose1d(x.reshape(-1, sp).unsqueeze(1), weight, stride=stride, padding=0)
    return up.squeeze(1).view(n, t, -1)


def fixed_stride_p_choose(p_choose_pooled: Tensor, src_len: int, ratio: int) -> Tensor:
    """Tail of FixedStrideMonotonicAttention.p_choose (modules/fixed_pre_decision.py:139-159):
    zero-upsample, then either append zeros (:141-153, upsampled row shorter than src_len -- the
    floor-pooled inference case) or cut to src_len and overwrite the last column with the last
    pooled value (:154-159)."""
    p = insert_zeros(p_choose_pooled, ratio)
    n, t, _ = p.shape
    if p.size(-1) < src_len:
        p = torch.cat([p, torch.zeros(n, t, src_len - p.size(-1)).to(p)], dim=2)
    else:
        p = p[:, :, :src_len]
        p[:, :, -1] = p_choose_pooled[:, :, -1]
    return p


def mma_process_train_pooled(p_choose_pooled: Tensor, src_len: int, ratio: int,
                             soft_energy: Optional[Tensor], padding_mask: Optional[Tensor] = None,
                             eps: float = 1e-6, mass_preserve: bool = True,
                             chunk_size: Optional[int] = None, compute_dtype: torch.dtype = F32):
    """monotonic_attention_process_train of a `*_fixed_pre_decision` class after the pooled
    p_choose_from_qk and the soft-energy bmm.  Returns (p_choose, alpha, beta)."""
    p = fixed_stride_p_choose(p_choose_pooled, src_len, ratio)
    alpha, beta = mma_process_train(p, soft_energy, padding_mask, eps, mass_preserve, chunk_size, compute_dtype)
    return p, alpha, beta


# --------------------------------------------------------------------------- a9
def mma_process_infer(
    p_choose: Tensor,            # [N, S]  (N = bsz * heads), sigmoid(energy), eval mode
    head_step: Tensor,           # [N] int64 carried state (zeros on the first call)
    soft_energy: Optional[Tensor] = None,   # [N, 1, S] (already padding-masked) or None
    padding_mask: Optional[Tensor] = None,  # [N, S] bool
    mass_preserve: bool = True,
):
    """Body of monotonic_attention_process_infer
    (modules/monotonic_multihead_attention.py:171-299) after the energy bmm's.

    Returns (new_step [N] int64, head_read [N] bool, alpha [N, S], beta [N, 1, S]).
    """
    n, s_len = p_choose.shape
    if padding_mask is not None:
        src_lengths = (~padding_mask).sum(1, keepdim=True)
    else:
        src_lengths = torch.full((n, 1), s_len, dtype=torch.long)
    assert int(src_lengths.max()) <= s_len
    step_in = head_step.view(n, 1)

    if mass_preserve:
        max_steps = src_lengths - 1
        work = p_choose.clone()
    else:
        max_steps = src_lengths
        work = torch.cat((p_choose, p_choose.new_zeros(n, 1)), dim=1)

    cols = torch.arange(work.shape[1]).view(1, -1)
    work = work.masked_fill(cols < step_in, 0.0)                 # :212-217 mask the past
    assert int(max_steps.max()) < work.shape[1]
    work = work.scatter(1, max_steps, 1.0)                       # :221-226 forced stop

    hit = work.ge(0.5)
    new_step = (hit.cumsum(1).eq(1)).int().argmax(1, keepdim=True)   # :230-237 first hit
    clamped = new_step.clamp(min=0)
    clamped = torch.minimum(clamped, src_lengths - 1)
    p_at = p_choose.gather(1, clamped)
    head_read = new_step.eq(max_steps) & (p_at < 0.5)            # :255-257

    alpha = torch.zeros_like(p_choose).scatter(1, clamped, 1)    # :261-268
    if not mass_preserve:
        alpha = alpha.masked_fill(new_step == max_steps, 0)      # :270-275

    if soft_energy is not None:                                   # :278-294
        beta_mask = torch.arange(s_len).expand_as(alpha).gt(new_step).unsqueeze(1)
        fill = -1e4 if soft_energy.dtype == torch.float16 else -1e8
        beta = torch.softmax(soft_energy.masked_fill(beta_mask, fill), dim=-1)
        beta = beta.masked_fill(new_step.eq(0).unsqueeze(1), 0)
    else:
        beta = alpha.view(n, 1, s_len)
    return new_step.view(n), head_read.view(n), alpha, beta


def expected_delays(alpha):
    """Step 2 of MMACriterion.compute_latency_loss (reference
    codebase/criterion/mma_criterion.py:146-157): `steps = arange(1, 1+src_len)` broadcast over
    alpha, `expected_delays = sum(steps * alpha, dim=-1)`.  The criterion module itself needs
    fairseq to import, so this three-line expression is restated, not loaded (parity unpinned by
    reference fixtures, like the rest of the latency loss -- SURVEY 8c)."""
    import torch
    src_len = alpha.size(-1)
    steps = torch.arange(1, 1 + src_len).unsqueeze(0).unsqueeze(1).expand_as(alpha).type_as(alpha)
    return torch.sum(steps * alpha, dim=-1)
