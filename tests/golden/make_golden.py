#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference functions are executed where they lie (oracle/ref_loader.py); only their
inputs and outputs are stored, as float32/int64 numpy arrays in compressed .npz files.
All inputs come from seeded CPU generators so the files are reproducible.

Files:
  mma_train.npz  -- expected_alignment_from_p_choose -> mass_preservation ->
                    expected_soft_attention (+ autograd gradients) on small cases
  mma_step.npz   -- monotonic_attention_process_infer over consecutive decoding steps
  cif.npz        -- cif_function (training and inference mode, + gradients)
  moving_sum.npz -- the hand-written example in utils/functions.py:83-105
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


# ----------------------------------------------------------------------------- MMA train
MMA_CASES = [
    # name, N, T, S, masked, chunk, soft, mass_preservation, energy_mean, sparse8
    ("il_nomask",      4,  6,  33, False, None, True,  True,  -2.0, False),
    ("il_mask",        3,  5,  64, True,  None, True,  True,  -2.0, False),
    ("chunk4_nomask",  3,  7,  40, False, 4,    True,  True,  -2.0, False),
    ("chunk3_mask",    3,  5,  48, True,  3,    True,  True,  -2.0, False),
    ("hard_nomp",      2,  4,  20, False, None, False, False, -2.0, False),
    ("hard_mp_mask",   2,  4,  24, True,  None, False, True,  -2.0, False),
    ("il_nomp",        2,  5,  31, False, None, True,  False, -2.0, False),
    ("il_c1rows",      2, 32, 256, False, None, True,  True,  -2.0, False),
    ("il_stress",      2, 16, 128, False, None, True,  True,   0.0, False),
    ("il_sparse8",     2,  8,  96, True,  None, True,  True,  -1.0, True),
    ("il_s1",          2,  3,   1, False, None, True,  True,  -2.0, False),
    ("il_odd",         3,  4,  67, True,  None, True,  True,  -2.0, False),
]


def gen_mma_train():
    _, ma, _ = ref_loader.load_utils()
    out = {}
    names = []
    for idx, (name, n, t, s, masked, chunk, soft, mp, mu, sparse8) in enumerate(MMA_CASES):
        g = torch.Generator().manual_seed(1000 + idx)
        energy = torch.randn(n, t, s, generator=g) + mu
        p = torch.sigmoid(energy)
        if sparse8:     # fixed pre-decision ratio 8: p is non-zero on every 8th column
            keep = (torch.arange(s) % 8) == 7
            p = p * keep
        soft_e = torch.randn(n, t, s, generator=g)
        mask = None
        if masked:
            lens = torch.randint(max(1, s // 2), s + 1, (n,), generator=g)
            lens[0] = s
            mask = torch.arange(s)[None, :] >= lens[:, None]
        g_alpha = torch.randn(n, t, s, generator=g)
        g_beta = torch.randn(n, t, s, generator=g)

        p = p.clone().requires_grad_()
        soft_e = soft_e.clone().requires_grad_()
        alpha = ma.expected_alignment_from_p_choose(p.float(), mask, eps=1e-6)
        if mp:
            alpha = ma.mass_preservation(alpha, mask)
        if soft:
            beta = ma.expected_soft_attention(alpha, soft_e, padding_mask=mask,
                                              chunk_size=chunk, eps=1e-6)
            loss = (alpha * g_alpha).sum() + (beta * g_beta).sum()
        else:
            beta = alpha
            loss = (alpha * g_alpha).sum()
        loss.backward()

        names.append(name)
        out[f"{name}/p"] = _np(p)
        out[f"{name}/soft_energy"] = _np(soft_e)
        out[f"{name}/mask"] = _np(mask) if mask is not None else np.zeros((0,), bool)
        out[f"{name}/g_alpha"] = _np(g_alpha)
        out[f"{name}/g_beta"] = _np(g_beta)
        out[f"{name}/alpha"] = _np(alpha)
        out[f"{name}/beta"] = _np(beta)
        out[f"{name}/grad_p"] = _np(p.grad)
        out[f"{name}/grad_soft_energy"] = (_np(soft_e.grad) if soft_e.grad is not None
                                           else np.zeros((0,), np.float32))
        out[f"{name}/cfg"] = np.array([n, t, s, int(masked), chunk or 0, int(soft), int(mp)],
                                      np.int64)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mma_train.npz"), **out)
    return names


# ----------------------------------------------------------------------------- MMA step
STEP_CASES = [
    # name, kind, mass_preservation, bsz, heads, S, steps, masked
    ("hard_mp",   "hard_aligned",      True,  5, 4, 37, 24, True),
    ("hard_nomp", "hard_aligned",      False, 5, 4, 37, 24, True),
    ("il_mp",     "infinite_lookback", True,  5, 4, 41, 24, True),
    ("il_nomp",   "infinite_lookback", False, 4, 2, 29, 20, False),
]


def gen_mma_step():
    out = {}
    names = []
    for idx, (name, kind, mp, bsz, heads, s, steps, masked) in enumerate(STEP_CASES):
        att = ref_loader.make_attention(kind, 8 * heads, heads, mass_preservation=mp).eval()
        g = torch.Generator().manual_seed(3000 + idx)
        n = bsz * heads
        # energies ~ N(-2, 1.5): heads advance a few frames per step, some stall
        p_all = torch.sigmoid(torch.randn(steps, n, s, generator=g) * 1.5 - 2.0)
        se_all = torch.randn(steps, n, s, generator=g)
        if masked:
            lens = torch.randint(max(2, s // 2), s + 1, (bsz,), generator=g)
            lens[0] = s
            kpm = torch.arange(s)[None, :] >= lens[:, None]
        else:
            lens = torch.full((bsz,), s)
            kpm = None
        inc = {}
        rec = {k: [] for k in ("head_step", "head_read", "alpha", "beta")}
        query = torch.zeros(1, bsz, 8 * heads)
        key = torch.zeros(s, bsz, 8 * heads)
        for st in range(steps):
            p_now = p_all[st]
            se_now = se_all[st]
            kpm_h = torch.repeat_interleave(kpm, heads, 0) if kpm is not None else None
            if kpm_h is not None:   # what energy_from_qk does to padded columns (:124-128)
                se_now = se_now.masked_fill(kpm_h, -1e8)
                p_now = torch.sigmoid(torch.logit(p_now).masked_fill(kpm_h, -1e8))
                p_all[st] = p_now
                se_all[st] = se_now
            att.p_choose = lambda q, k, m, inc_state=None, _p=p_now: _p.unsqueeze(1)
            att.energy_from_qk = (lambda q, k, t, key_padding_mask=None, bias=0, _e=se_now:
                                  _e.unsqueeze(1))
            with torch.no_grad():
                _, alpha, beta = att.monotonic_attention_process_infer(query, key, kpm_h, inc)
            cache = att._get_monotonic_buffer(inc)
            rec["head_step"].append(cache["head_step"].reshape(n).clone())
            rec["head_read"].append(cache["head_read"].reshape(n).clone())
            rec["alpha"].append(alpha.reshape(n, s).clone())
            rec["beta"].append(beta.reshape(n, s).clone())
        names.append(name)
        out[f"{name}/p"] = _np(p_all)
        out[f"{name}/soft_energy"] = _np(se_all)
        out[f"{name}/src_lengths"] = _np(torch.repeat_interleave(lens, heads, 0))
        out[f"{name}/cfg"] = np.array([int(kind != "hard_aligned"), int(mp), bsz, heads, s,
                                       steps, int(masked)], np.int64)
        for k, v in rec.items():
            out[f"{name}/{k}"] = _np(torch.stack(v))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "mma_step.npz"), **out)
    return names


# ----------------------------------------------------------------------------- CIF
CIF_CASES = [
    # name, B, S, C, beta, masked, train, alpha_mean
    ("train_b1",      4,  50,  6, 1.0,  True,  True,  -1.0),
    ("infer_b1",      4,  50,  6, 1.0,  True,  False, -1.0),
    ("train_multi",   3,  40,  5, 0.35, True,  True,   0.5),
    ("infer_multi",   3,  40,  5, 0.35, False, False,  0.5),
    ("train_b13",     3,  64,  8, 1.3,  False, True,   0.0),
    ("infer_b13",     3,  64,  8, 1.3,  True,  False,  0.0),
    ("infer_short",   2,   3,  4, 1.0,  False, False, -2.0),   # nothing fires
    ("train_s1",      2,   1,  3, 1.0,  False, True,   0.0),
    ("infer_c256",    2, 120, 256, 1.0, True,  False, -1.0),
    ("train_c256",    2, 120, 256, 1.0, True,  True,  -1.0),
]


def gen_cif():
    rc = ref_loader.load_cif()
    out = {}
    names = []
    for idx, (name, b, s, c, beta, masked, train, mu) in enumerate(CIF_CASES):
        g = torch.Generator().manual_seed(2000 + idx)
        x = torch.randn(b, s, c, generator=g).requires_grad_()
        alpha = torch.sigmoid(torch.randn(b, s, generator=g) + mu).requires_grad_()
        mask = None
        if masked:
            lens = torch.randint(max(1, s // 2), s + 1, (b,), generator=g)
            lens[0] = s
            mask = torch.arange(s)[None, :] >= lens[:, None]
        kw = {}
        if train:
            am = alpha.detach() if mask is None else alpha.detach().masked_fill(mask, 0)
            tl = (am.sum(1) / beta).round().clamp(min=1).long()
            tl[-1] = max(1, int(tl[-1]) - 2)      # a row shorter than T
            kw["target_lengths"] = tl
        res = rc.cif_function(x, alpha, beta=beta, tail_thres=beta / 2, padding_mask=mask, **kw)
        cif_out, delays = res["cif_out"][0], res["delays"][0]
        g_out = torch.randn(cif_out.shape, generator=g)
        g_delay = torch.randn(delays.shape, generator=g)
        ((cif_out * g_out).sum() + (delays * g_delay).sum()).backward()
        names.append(name)
        out[f"{name}/input"] = _np(x)
        out[f"{name}/alpha"] = _np(alpha)
        out[f"{name}/mask"] = _np(mask) if mask is not None else np.zeros((0,), bool)
        out[f"{name}/target_lengths"] = (_np(kw["target_lengths"]) if train
                                         else np.zeros((0,), np.int64))
        out[f"{name}/cfg"] = np.array([b, s, c, int(masked), int(train)], np.int64)
        out[f"{name}/beta"] = np.array(beta, np.float64)
        out[f"{name}/cif_out"] = _np(cif_out)
        out[f"{name}/cif_lengths"] = _np(res["cif_lengths"][0])
        out[f"{name}/alpha_sum"] = _np(res["alpha_sum"][0])
        out[f"{name}/delays"] = _np(delays)
        out[f"{name}/tail_weights"] = (_np(res["tail_weights"][0]) if not train
                                       else np.zeros((0,), np.float32))
        out[f"{name}/g_out"] = _np(g_out)
        out[f"{name}/g_delay"] = _np(g_delay)
        out[f"{name}/grad_input"] = _np(x.grad)
        out[f"{name}/grad_alpha"] = _np(alpha.grad)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "cif.npz"), **out)
    return names


def gen_moving_sum():
    fn, _, _ = ref_loader.load_utils()
    # docstring example (functions.py:83-105); its columns are this tensor's last axis
    x = torch.arange(15.0).view(3, 5).unsqueeze(0)
    np.savez_compressed(
        os.path.join(HERE, "moving_sum.npz"),
        x=_np(x), s3e1=_np(fn.moving_sum(x, 3, 1)), s1e3=_np(fn.moving_sum(x, 1, 3)),
        doc_s3e1=np.array([[0, 1, 3, 6, 9], [5, 11, 18, 21, 24], [10, 21, 33, 36, 39]], np.float32),
        doc_s1e3=np.array([[3, 6, 9, 7, 4], [18, 21, 24, 17, 9], [33, 36, 39, 27, 14]], np.float32),
    )


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not found at " + ref_loader.REF_ROOT)
    torch.set_num_threads(1)
    print("mma_train:", gen_mma_train())
    print("mma_step :", gen_mma_step())
    print("cif      :", gen_cif())
    gen_moving_sum()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
