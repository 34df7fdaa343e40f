"""CPU: oracle restatements and torch-only host mirrors against the round-2 golden vectors
(tests/golden/make_golden_r2.py, generated from the unmodified reference)."""
import pytest
import torch

from oracle import mma as omma
from tests.golden_io import load, opt

WAITK = load("waitk.npz")
LEFT = load("mma_leftpad.npz")
FIXED = load("fixed_predecision.npz")


@pytest.mark.parametrize("name", list(WAITK))
def test_waitk_p_choose_mirror_and_oracle_match_golden(name):
    """a11: the mirror generates only the last target row (the reference's unconditional
    ``[:, -1:]``); result, dtype and shape must equal the reference's."""
    from simulst_b200.utils.p_choose_strategy import waitk_p_choose
    c = WAITK[name]
    t, s, b, k, masked, online = [int(v) for v in c.cfg]
    inc = {"online": True} if online else {}
    got = waitk_p_choose(t, s, b, k, opt(c.mask), inc)
    assert got.dtype == torch.bool and tuple(got.shape) == tuple(c.p_choose.shape) == (b, 1, s)
    assert torch.equal(got, c.p_choose)
    ora = omma.waitk_p_choose(t, s, b, k, opt(c.mask), online=bool(online), last_only=True)
    assert torch.equal(ora, c.p_choose)


def test_waitk_p_choose_without_incremental_state_raises_like_upstream():
    """p_choose_strategy.py:35 dereferences incremental_state unconditionally."""
    from simulst_b200.utils.p_choose_strategy import waitk_p_choose
    with pytest.raises(AttributeError):
        waitk_p_choose(3, 4, 1, 1, None, None)


def test_left_padding_oracle_matches_golden():
    c = LEFT["left"]
    p = c.p.clone().requires_grad_()
    a = omma.expected_alignment_from_p_choose(p, c.mask, eps=1e-6)
    a = omma.mass_preservation(a, c.mask, left_padding=True)
    (a * c.g_alpha).sum().backward()
    torch.testing.assert_close(a, c.alpha, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(p.grad, c.grad_p, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(FIXED))
def test_fixed_pre_decision_oracle_matches_golden(name):
    """8f #2: insert_zeros + tail fix-up + the training path, restated, against what the reference's
    own *_fixed_pre_decision classes produced (outputs bit-identical, autograd gradients of the
    POOLED p_choose and of the soft energy)."""
    c = FIXED[name]
    n, t, s, ratio, masked, soft, mp = [int(v) for v in c.cfg]
    pp = c.p_pooled.clone().requires_grad_()
    se = c.soft_energy.clone().requires_grad_() if soft else None
    p, alpha, beta = omma.mma_process_train_pooled(pp, s, ratio, se, opt(c.mask), 1e-6, bool(mp))
    assert torch.equal(p, c.p_choose)
    loss = (alpha * c.g_alpha).sum()
    if soft:
        loss = loss + (beta * c.g_beta).sum()
    loss.backward()
    assert torch.equal(alpha, c.alpha) and torch.equal(beta, c.beta)
    torch.testing.assert_close(pp.grad, c.grad_p_pooled, rtol=1e-6, atol=1e-7)
    if soft:
        torch.testing.assert_close(se.grad, c.grad_soft_energy, rtol=1e-6, atol=1e-7)
