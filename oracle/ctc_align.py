"""Oracle (CPU numpy / PyTorch) for the CTC best alignment -- SURVEY 8f rank 3.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Reference (the only native code the reference ships, CUDA-only, JIT-built at import):
  codebase/criterion/best_alignment/best_alignment.cu:58-202   ``ctc_alignment_log_alpha_gpu_kernel``
      Viterbi ("max" instead of log-sum-exp) forward over the blank-augmented target
      l' = blank l_0 blank l_1 ... blank (2T+1 states), recording the arg-max predecessor
  codebase/criterion/best_alignment/best_alignment.cu:204-313  host template (shapes, -1 filled paths)
  codebase/criterion/best_alignment/__init__.py:25-111         final-state choice, S-step back-trace,
                                                               optional state -> label translation

``viterbi_forward`` restates the kernel; ``best_alignment`` restates the Python wrapper.  Pinning:
  * the wrapper restatement is checked against the reference's OWN wrapper source executed over
    ``viterbi_forward`` (ref_loader.load_best_alignment_python), here on CPU;
  * the kernel restatement has no CPU reference to run against (the reference kernel is CUDA-only),
    so it is pinned (a) by an exhaustive search over every valid CTC path on tiny cases
    (``bruteforce_best_path``: the definition of "best alignment"), here on CPU, and (b) on the GPU
    box by the reference's own kernel, JIT-built from baseline/_ref exactly like the reference does
    (ref_loader.load_best_alignment_extension), in tests/test_ctc_align.py.
"""
import itertools
from types import SimpleNamespace

import numpy as np
import torch

NEG_INF = float("-inf")


def _augmented(target_row, tl, blank):
    """l' of Graves et al. (best_alignment.cu:33-42): blank at even positions."""
    aug = np.full(2 * tl + 1, blank, dtype=np.int64)
    aug[1::2] = target_row[:tl]
    return aug


def viterbi_forward(log_prob, targets, input_lengths, target_lengths, blank=0):
    """best_alignment.cu:58-202 + :287-313.  log_prob (S, N, V); targets (N, Tmax).
    Returns (nll (N,), log_alpha (N, S, 2*Tmax+1), paths (N, S, 2*Tmax+1) int64, -1 where unset)."""
    lp = log_prob.detach().cpu().numpy().astype(np.float64 if log_prob.dtype == torch.float64 else np.float32)
    tg = targets.detach().cpu().numpy()
    il = [int(v) for v in input_lengths]
    tl = [int(v) for v in target_lengths]
    s_len, n, _ = lp.shape
    t_max = max(tl) if tl else 0
    width = 2 * t_max + 1
    la = np.empty((n, s_len, width), dtype=lp.dtype)            # at::empty: every cell is written below
    paths = np.full((n, s_len, width), -1, dtype=np.int64)
    nll = np.zeros(n, dtype=lp.dtype)
    for b in range(n):
        aug = _augmented(tg[b], tl[b], blank)
        states = 2 * tl[b] + 1
        la[b, 0, :] = NEG_INF
        la[b, 0, 0] = lp[0, b, blank]
        if tl[b] > 0 and width > 1:
            la[b, 0, 1] = lp[0, b, aug[1]]
        three = np.zeros(states, dtype=bool)
        three[2:] = aug[2:] != aug[:-2]
        for t in range(1, s_len):
            la[b, t, :] = NEG_INF
            if t >= il[b]:
                continue
            prev = la[b, t - 1, :states]
            best = prev.copy()
            arg = np.arange(states)
            stay1 = np.concatenate(([NEG_INF], prev[:-1]))
            take = stay1 > best                                 # strict: ties keep the smaller jump
            best = np.where(take, stay1, best)
            arg = np.where(take, np.arange(states) - 1, arg)
            stay2 = np.concatenate(([NEG_INF, NEG_INF], prev[:-2]))[:states]
            take = three & (stay2 > best)
            best = np.where(take, stay2, best)
            arg = np.where(take, np.arange(states) - 2, arg)
            la[b, t, :states] = best + lp[t, b, aug]
            paths[b, t, :states] = arg
        l1 = la[b, il[b] - 1, 2 * tl[b]]
        l2 = la[b, il[b] - 1, 2 * tl[b] - 1] if tl[b] > 0 else NEG_INF
        m = max(l1, l2)
        m = 0.0 if m == NEG_INF else m
        with np.errstate(divide="ignore"):
            nll[b] = -(np.log(np.exp(l1 - m) + np.exp(l2 - m)) + m)
    return torch.from_numpy(nll), torch.from_numpy(la), torch.from_numpy(paths)


def as_extension():
    """An object with the reference extension's call shape (`extension.best_alignment(...)`,
    best_alignment.cpp:10-31) backed by ``viterbi_forward``."""
    def best_alignment(log_prob, targets, input_lengths, target_lengths, blank, zero_infinity):
        return viterbi_forward(log_prob, targets, input_lengths, target_lengths, blank)
    return SimpleNamespace(best_alignment=best_alignment)


def best_alignment(log_prob, targets, input_lengths, target_lengths, blank=0, as_labels=False):
    """best_alignment/__init__.py:25-111 restated: states (N, S) int64 (or labels)."""
    _, la, paths = viterbi_forward(log_prob, targets, input_lengths, target_lengths, blank)
    n, s_len, width = la.shape
    out = torch.zeros(n, s_len, dtype=torch.long)
    for b in range(n):
        il, tl = int(input_lengths[b]), int(target_lengths[b])
        states = 2 * tl + 1
        end = la[b, il - 1]
        neg = (end == NEG_INF).nonzero()
        first_neg = int(neg[0]) if neg.numel() else 0           # argmax of an all-zero row is 0
        last = (first_neg - 1) % states
        last = min(last, states - 2)
        cand = end.clone()
        idx = torch.arange(width)
        cand[(idx < last) | (idx >= states)] = NEG_INF
        cur = int(torch.argmax(cand))                           # first maximum
        out[b, il - 1] = cur
        for t in range(il - 1, 0, -1):
            cur = int(paths[b, t, cur])
            out[b, t - 1] = cur
        # frames t >= input_length: arg-max over an all -inf column = 0 (already zero)
    if as_labels:
        lab = targets.gather(1, out.div(2, rounding_mode="floor").clamp(max=max(targets.shape[1] - 1, 0)))
        return torch.where(out % 2 == 1, lab, torch.full_like(out, blank))
    return out


def bruteforce_best_path(log_prob, target, blank=0):
    """Definition of the best alignment for ONE sample: the maximum-probability state sequence
    through l' that starts in state 0 or 1, ends in one of the last two states, moves by 0, +1, or
    +2 (the latter only between different non-blank labels).  Exhaustive; tiny cases only.
    Returns (best log-probability, state list)."""
    lp = log_prob.detach().cpu().double().numpy()
    s_len = lp.shape[0]
    aug = _augmented(np.asarray(target), len(target), blank)
    states = len(aug)
    best, best_path = NEG_INF, None
    for path in itertools.product(range(states), repeat=s_len):
        if path[0] > 1 or path[-1] < states - 2:
            continue
        ok = True
        for a, b in zip(path[:-1], path[1:]):
            d = b - a
            if d < 0 or d > 2 or (d == 2 and not (b % 2 == 1 and aug[b] != aug[a])):
                ok = False
                break
        if not ok:
            continue
        score = sum(lp[t, aug[s]] for t, s in enumerate(path))
        if score > best:
            best, best_path = score, list(path)
    return best, best_path
