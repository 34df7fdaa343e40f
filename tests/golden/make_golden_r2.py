#!/usr/bin/env python
"""Round-2 golden vectors, generated from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_r2.py

Kept apart from make_golden.py so the round-1 files stay byte-identical.

Files:
  waitk.npz          -- waitk_p_choose (utils/p_choose_strategy.py:6-53)
  latency.npz        -- MMACriterion.compute_latency_loss (criterion/mma_criterion.py:138-207),
                        the method's own source executed with SimulEval's DAL restated
                        (oracle/latency.py) as its LATENCY_METRICS
  mma_leftpad.npz    -- expected_alignment + mass_preservation(left_padding=True)
  mma_module.npz     -- MonotonicAttention / MonotonicInfiniteLookbackAttention .forward()
                        (modules/monotonic_multihead_attention.py:354-423): training pass with
                        gradients of every parameter, and 16 incremental decoding steps
  fixed_predecision.npz -- the *_fixed_pre_decision wrappers (modules/fixed_pre_decision.py):
                        pooled p_choose -> insert_zeros -> tail fix-up -> alpha / beta, + grads
  ssnt.npz           -- ssnt_loss / ssnt_loss_mem (criterion/ssnt_loss/ssnt_loss.py:45-271)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import latency as olat  # noqa: E402
from oracle import ref_loader  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _save(file_name, out, names):
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, file_name), **out)
    return names


# ----------------------------------------------------------------------------- wait-k
WAITK_CASES = [
    # name, tgt_len, src_len, bsz, k, masked, online
    ("k3", 5, 9, 3, 3, False, False),
    ("k1_mask", 7, 6, 2, 1, True, False),
    ("k5_online", 4, 20, 3, 5, True, True),
    ("one", 1, 1, 1, 1, False, False),
    ("beyond_online", 12, 6, 2, 3, False, True),
    ("beyond_clip", 12, 6, 2, 3, True, False),
]


def gen_waitk():
    _, _, pc = ref_loader.load_utils()
    out, names = {}, []
    for idx, (name, t, s, b, k, masked, online) in enumerate(WAITK_CASES):
        g = torch.Generator().manual_seed(4000 + idx)
        mask = None
        if masked:
            lens = torch.randint(1, s + 1, (b,), generator=g)
            mask = torch.arange(s)[None, :] >= lens[:, None]
        inc = {"online": True} if online else {}
        res = pc.waitk_p_choose(t, s, b, k, mask, inc)
        names.append(name)
        out[f"{name}/cfg"] = np.array([t, s, b, k, int(masked), int(online)], np.int64)
        out[f"{name}/mask"] = _np(mask) if mask is not None else np.zeros((0,), bool)
        out[f"{name}/p_choose"] = _np(res)
    return _save("waitk.npz", out, names)


# ----------------------------------------------------------------------------- latency loss
LATENCY_CASES = [
    # name, bsz, layers, heads, T, S, gather, avg_w, var_w
    ("weighted", 3, 2, 2, 7, 40, "weighted_average", 0.1, 0.1),
    ("max", 2, 3, 2, 9, 33, "max", 0.5, 0.0),
    ("weighted_long", 2, 1, 4, 24, 96, "weighted_average", 1.0, 0.3),
]


def gen_latency():
    _, ma, _ = ref_loader.load_utils()
    fn = ref_loader.load_mma_latency_loss(olat.LATENCY_METRICS)
    out, names = {}, []
    for idx, (name, bsz, layers, heads, t, s, gather, avg_w, var_w) in enumerate(LATENCY_CASES):
        g = torch.Generator().manual_seed(5000 + idx)
        enc_len = torch.randint(max(2, s // 2), s + 1, (bsz,), generator=g)
        enc_len[0] = s
        enc_mask = torch.arange(s)[None, :] >= enc_len[:, None]
        tgt_len = torch.randint(max(1, t // 2), t + 1, (bsz,), generator=g)
        tgt_len[0] = t
        pad = 1
        target = torch.randint(2, 50, (bsz, t), generator=g)
        target[torch.arange(t)[None, :] >= tgt_len[:, None]] = pad
        src_lengths = enc_len * 4 + torch.randint(0, 4, (bsz,), generator=g)     # frames before subsampling
        p_list, alpha_list = [], []
        for _ in range(layers):
            p = torch.sigmoid(torch.randn(bsz * heads, t, s, generator=g) - 2.0).requires_grad_()
            mask_h = torch.repeat_interleave(enc_mask, heads, 0)
            alpha = ma.expected_alignment_from_p_choose(p.float(), mask_h, eps=1e-6)
            alpha = ma.mass_preservation(alpha, mask_h)
            p_list.append(p)
            alpha_list.append(alpha.view(bsz, heads, t, s))
        cfg = olat.criterion_stub(avg_w, var_w, gather, pad, 10.0)
        sample = {"target": target, "net_input": {"src_lengths": src_lengths}}
        net_output = (None, {"attn_list": [{"alpha": a} for a in alpha_list],
                             "encoder_padding_mask": [enc_mask]})
        loss, latency, var = fn(cfg, None, sample, net_output)
        loss.backward()
        steps = torch.arange(1, 1 + s).float()
        delays = torch.cat(alpha_list, dim=1).view(-1, t, s).detach().mul(steps).sum(-1)
        names.append(name)
        out[f"{name}/cfg"] = np.array([bsz, layers, heads, t, s], np.int64)
        out[f"{name}/weights"] = np.array([avg_w, var_w], np.float64)
        out[f"{name}/gather"] = np.array(gather)
        out[f"{name}/p"] = _np(torch.stack(p_list))                     # [layers, bsz*heads, T, S]
        out[f"{name}/grad_p"] = _np(torch.stack([p.grad for p in p_list]))
        out[f"{name}/enc_mask"] = _np(enc_mask)
        out[f"{name}/target"] = _np(target)
        out[f"{name}/src_lengths"] = _np(src_lengths)
        out[f"{name}/expected_delays"] = _np(delays)                    # [bsz*layers*heads, T]
        out[f"{name}/latency_loss"] = _np(loss)
        out[f"{name}/expected_latency"] = _np(latency)
        out[f"{name}/delays_var"] = _np(var)
    return _save("latency.npz", out, names)


# ----------------------------------------------------------------------------- left padding
def gen_leftpad():
    _, ma, _ = ref_loader.load_utils()
    out, names = {}, []
    n, t, s = 4, 6, 96
    g = torch.Generator().manual_seed(23)
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2.0).requires_grad_()
    lens = torch.tensor([96, 70, 51, 96])
    mask = torch.arange(s)[None, :] < (s - lens)[:, None]
    ga = torch.randn(n, t, s, generator=g)
    alpha = ma.expected_alignment_from_p_choose(p, mask, eps=1e-6)
    alpha = ma.mass_preservation(alpha, mask, left_padding=True)
    (alpha * ga).sum().backward()
    name = "left"
    names.append(name)
    out[f"{name}/p"] = _np(p)
    out[f"{name}/mask"] = _np(mask)
    out[f"{name}/g_alpha"] = _np(ga)
    out[f"{name}/alpha"] = _np(alpha)
    out[f"{name}/grad_p"] = _np(p.grad)
    return _save("mma_leftpad.npz", out, names)


# ----------------------------------------------------------------------------- SSNT
SSNT_CASES = [
    # name, N, T, S, V, use_logits, ragged, fastemit_lambda, reduction
    ("logits_full",   3, 5, 20, 7,  True,  False, 0.0,  "none"),
    ("probs_full",    3, 5, 20, 7,  False, False, 0.0,  "none"),
    ("logits_ragged", 4, 7, 33, 11, True,  True,  0.0,  "sum"),
    ("probs_ragged",  4, 7, 33, 11, False, True,  0.0,  "mean"),
    ("fastemit",      2, 6, 40, 5,  True,  True,  0.01, "sum"),
    ("long",          2, 24, 150, 9, True, True,  0.0,  "sum"),
    ("one",           1, 1, 1, 3,   True,  False, 0.0,  "none"),
]


def gen_ssnt():
    ref = ref_loader.load_ssnt()
    out, names = {}, []
    for idx, (name, n, t, s, v, use_logits, ragged, lam, red) in enumerate(SSNT_CASES):
        g = torch.Generator().manual_seed(6000 + idx)
        logits = torch.randn(n, t, s, v, generator=g).requires_grad_()
        emit = (torch.randn(n, t, s, generator=g) - 1.0)
        targets = torch.randint(0, v, (n, t), generator=g)
        if ragged:
            src_len = torch.randint(max(1, s // 2), s + 1, (n,), generator=g)
            tgt_len = torch.randint(max(1, t // 2), t + 1, (n,), generator=g)
            src_len[0], tgt_len[0] = s, t
        else:
            src_len, tgt_len = torch.full((n,), s), torch.full((n,), t)
        if use_logits:
            emit_in = emit.clone().requires_grad_()
            kw = {"emit_logits": emit_in}
        else:
            emit_in = torch.sigmoid(emit).clone().requires_grad_()
            kw = {"emit_probs": emit_in}
        lp = logits.log_softmax(-1)
        loss, lattice, log_p = ref.ssnt_loss(lp, targets, src_len, tgt_len, reduction=red,
                                             fastemit_lambda=lam, **kw)
        w = torch.randn(loss.shape, generator=g) if loss.dim() else torch.tensor(1.0)
        (loss * w).sum().backward()
        # the memory-efficient variant on the same sample (targets concatenated)
        keep = torch.arange(t)[None, :] < tgt_len[:, None]
        kw_m = {k: (x.detach()[keep]) for k, x in kw.items()}
        loss_m, lattice_m, _ = ref.ssnt_loss_mem(lp.detach()[keep], targets[keep], src_len, tgt_len,
                                                 reduction=red, fastemit_lambda=lam, **kw_m)
        names.append(name)
        out[f"{name}/cfg"] = np.array([n, t, s, v, int(use_logits)], np.int64)
        out[f"{name}/fastemit"] = np.array(lam, np.float64)
        out[f"{name}/reduction"] = np.array(red)
        out[f"{name}/logits"] = _np(logits)
        out[f"{name}/emit"] = _np(emit_in)
        out[f"{name}/targets"] = _np(targets)
        out[f"{name}/source_lengths"] = _np(src_len)
        out[f"{name}/target_lengths"] = _np(tgt_len)
        out[f"{name}/w"] = _np(w)
        out[f"{name}/loss"] = _np(loss)
        out[f"{name}/lattice"] = _np(lattice)
        out[f"{name}/log_p_choose"] = _np(log_p)
        out[f"{name}/grad_logits"] = _np(logits.grad)
        out[f"{name}/grad_emit"] = _np(emit_in.grad)
        out[f"{name}/loss_mem"] = _np(loss_m)
        out[f"{name}/lattice_mem"] = _np(lattice_m)
    return _save("ssnt.npz", out, names)


# ----------------------------------------------------------------------------- fixed pre-decision
FIXED_CASES = [
    # name, kind, ratio, pool, bsz, heads, T, S, masked, mass_preservation
    ("il_r8_div",      "infinite_lookback", 8, "average", 2, 2, 6, 64,  False, True),
    ("il_r8_tail",     "infinite_lookback", 8, "average", 2, 2, 5, 61,  False, True),
    ("il_r8_mask",     "infinite_lookback", 8, "average", 3, 2, 5, 72,  True,  True),
    ("il_r8_tailmask", "infinite_lookback", 8, "last",    3, 2, 4, 75,  True,  True),
    ("hard_r4_tail",   "hard_aligned",      4, "last",    2, 2, 5, 30,  False, True),
    ("hard_r8_mask",   "hard_aligned",      8, "average", 2, 4, 4, 48,  True,  False),
    ("il_r8_short",    "infinite_lookback", 8, "average", 2, 2, 3, 5,   False, True),
    ("il_r3_odd",      "infinite_lookback", 3, "average", 2, 2, 4, 32,  False, True),
    ("il_r16_long",    "infinite_lookback", 16, "average", 1, 2, 8, 264, True, True),
]


def gen_fixed_predecision():
    """The real wrapper classes (modules/fixed_pre_decision.py:175-190) run through their own
    monotonic_attention_process_train; p_choose_from_qk and energy_from_qk are wrapped on the
    instance only to RECORD the pooled p_choose and the soft energy they return (the two tensors
    the kernels take), with retain_grad so their autograd gradients are stored too."""
    out, names = {}, []
    for idx, (name, kind, ratio, pool, bsz, heads, t, s, masked, mp) in enumerate(FIXED_CASES):
        embed = 8 * heads
        att = ref_loader.make_fixed_pre_decision_attention(kind, ratio, pool, embed, heads,
                                                           mass_preservation=mp, seed=7000 + idx)
        att.train()
        att.noise_std = 0.0
        g = torch.Generator().manual_seed(7100 + idx)
        q = torch.randn(t, bsz, embed, generator=g)
        k = torch.randn(s, bsz, embed, generator=g) * 1.5
        mask = None
        if masked:
            lens = torch.randint(max(2, s // 2), s + 1, (bsz,), generator=g)
            lens[0] = s
            mask = (torch.arange(s)[None, :] >= lens[:, None]).repeat_interleave(heads, 0)
        rec = {}
        orig_p, orig_e = att.p_choose_from_qk, att.energy_from_qk

        def rec_p(*a, **kw):
            v = orig_p(*a, **kw)
            v.retain_grad()
            rec["p_pooled"] = v
            return v

        def rec_e(query, key, energy_type, **kw):
            v = orig_e(query, key, energy_type, **kw)
            if energy_type == "soft":
                v.retain_grad()
                rec["soft_energy"] = v
            return v
        att.p_choose_from_qk, att.energy_from_qk = rec_p, rec_e
        p_choose, alpha, beta, _ = att.monotonic_attention_process_train(q, k, mask)
        soft = kind != "hard_aligned"
        n = bsz * heads
        g_alpha = torch.randn(n, t, s, generator=g)
        g_beta = torch.randn(n, t, s, generator=g)
        loss = (alpha * g_alpha).sum()
        if soft:
            loss = loss + (beta * g_beta).sum()
        loss.backward()
        names.append(name)
        out[f"{name}/cfg"] = np.array([n, t, s, ratio, int(masked), int(soft), int(mp)], np.int64)
        out[f"{name}/p_pooled"] = _np(rec["p_pooled"])
        out[f"{name}/grad_p_pooled"] = _np(rec["p_pooled"].grad)
        out[f"{name}/p_choose"] = _np(p_choose)
        out[f"{name}/soft_energy"] = _np(rec["soft_energy"]) if soft else np.zeros((0,), np.float32)
        out[f"{name}/grad_soft_energy"] = _np(rec["soft_energy"].grad) if soft else np.zeros((0,), np.float32)
        out[f"{name}/mask"] = _np(mask) if mask is not None else np.zeros((0,), bool)
        out[f"{name}/g_alpha"] = _np(g_alpha)
        out[f"{name}/g_beta"] = _np(g_beta)
        out[f"{name}/alpha"] = _np(alpha)
        out[f"{name}/beta"] = _np(beta)
    return _save("fixed_predecision.npz", out, names)


GENERATORS = [("waitk", gen_waitk), ("latency", gen_latency), ("mma_leftpad", gen_leftpad),
              ("ssnt", gen_ssnt), ("fixed_predecision", gen_fixed_predecision)]


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not found at " + ref_loader.REF_ROOT)
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    for name, fn in GENERATORS:
        if only and name not in only:
            continue
        print(f"{name:18s}", fn())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
