"""CPU: the oracle restatement must reproduce the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py), forward and gradients, bit for bit
in fp32 (same primitive sequence on the same CPU backend => tolerance is 0 here, with
a tiny atol for cross-machine libm differences)."""
import numpy as np
import pytest
import torch

from oracle import cif as ocif
from oracle import mma as omma
from tests.golden_io import load, opt

TRAIN = load("mma_train.npz")
STEP = load("mma_step.npz")
CIF = load("cif.npz")
TIGHT = dict(rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", list(TRAIN))
def test_mma_train_oracle_matches_golden(name):
    c = TRAIN[name]
    n, t, s, masked, chunk, soft, mp = [int(v) for v in c.cfg]
    p = c.p.clone().requires_grad_()
    se = c.soft_energy.clone().requires_grad_()
    alpha, beta = omma.mma_process_train(p, se if soft else None, opt(c.mask), 1e-6,
                                         bool(mp), chunk or None)
    torch.testing.assert_close(alpha, c.alpha, **TIGHT)
    torch.testing.assert_close(beta, c.beta, **TIGHT)
    loss = (alpha * c.g_alpha).sum()
    if soft:
        loss = loss + (beta * c.g_beta).sum()
    loss.backward()
    torch.testing.assert_close(p.grad, c.grad_p, rtol=1e-5, atol=1e-6)
    if soft:
        torch.testing.assert_close(se.grad, c.grad_soft_energy, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(TRAIN))
def test_mma_train_fp64_restatement_close_to_reference(name):
    """The fp64 evaluation of the same formulas bounds the reference's own fp32 error."""
    c = TRAIN[name]
    n, t, s, masked, chunk, soft, mp = [int(v) for v in c.cfg]
    alpha, beta = omma.mma_process_train(c.p, c.soft_energy if soft else None, opt(c.mask),
                                         1e-6, bool(mp), chunk or None,
                                         compute_dtype=torch.float64)
    torch.testing.assert_close(alpha.float(), c.alpha, rtol=2e-4, atol=2e-6)
    torch.testing.assert_close(beta.float(), c.beta, rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("name", list(STEP))
def test_mma_step_oracle_matches_golden(name):
    c = STEP[name]
    soft, mp, bsz, heads, s, steps, masked = [int(v) for v in c.cfg]
    n = bsz * heads
    mask = (torch.arange(s)[None, :] >= c.src_lengths[:, None]) if masked else None
    head_step = torch.zeros(n, dtype=torch.long)
    for st in range(steps):
        se = c.soft_energy[st].unsqueeze(1) if soft else None
        head_step, head_read, alpha, beta = omma.mma_process_infer(
            c.p[st], head_step, se, mask, bool(mp))
        assert torch.equal(head_step, c.head_step[st])
        assert torch.equal(head_read, c.head_read[st])
        assert torch.equal(alpha, c.alpha[st])
        torch.testing.assert_close(beta.reshape(n, s), c.beta[st], **TIGHT)


@pytest.mark.parametrize("name", list(CIF))
def test_cif_oracle_matches_golden(name):
    c = CIF[name]
    b, s, ch, masked, train = [int(v) for v in c.cfg]
    x = c.input.clone().requires_grad_()
    a = c.alpha.clone().requires_grad_()
    res = ocif.cif_function(x, a, beta=float(c.beta), tail_thres=float(c.beta) / 2,
                            padding_mask=opt(c.mask), target_lengths=opt(c.target_lengths))
    assert torch.equal(res["cif_lengths"][0], c.cif_lengths)
    torch.testing.assert_close(res["cif_out"][0], c.cif_out, **TIGHT)
    torch.testing.assert_close(res["delays"][0], c.delays, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(res["alpha_sum"][0], c.alpha_sum, **TIGHT)
    if not train:
        torch.testing.assert_close(res["tail_weights"][0], c.tail_weights, **TIGHT)
    ((res["cif_out"][0] * c.g_out).sum() + (res["delays"][0] * c.g_delay).sum()).backward()
    torch.testing.assert_close(x.grad, c.grad_input, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a.grad, c.grad_alpha, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", list(CIF))
def test_cif_golden_against_sequential_checker(name):
    """The reference's own acceptance test: parallel CIF vs the frame-by-frame loop at 1e-3
    (torch_cif/test.py:127-184)."""
    c = CIF[name]
    out, delay = ocif.cif_sequential(c.input, c.alpha, beta=float(c.beta),
                                     tail_thres=float(c.beta) / 2, padding_mask=opt(c.mask),
                                     target_lengths=opt(c.target_lengths))
    t = c.cif_out.shape[1]
    if out.shape[1] == t + 1:       # checker kept an all-but-zero extra slot
        out, delay = out[:, :t], delay[:, :t]
    torch.testing.assert_close(out.float()[:, :t], c.cif_out, rtol=1e-3, atol=1e-3)
    # delays: the reference only compares them where its checker slices identically
    if int(c.cfg[4]):
        torch.testing.assert_close(delay.float()[:, :t], c.delays, rtol=1e-3, atol=1e-3)


def test_moving_sum_docstring_example():
    z = np.load(__import__("os").path.join(__import__("tests.golden_io").golden_io.GOLDEN_DIR,
                                           "moving_sum.npz"))
    x = torch.from_numpy(z["x"])
    for key, (a, b) in {"s3e1": (3, 1), "s1e3": (1, 3)}.items():
        got = omma.moving_sum(x, a, b)[0]
        assert torch.equal(got, torch.from_numpy(z["doc_" + key]))
        assert torch.equal(got, torch.from_numpy(z[key])[0])
