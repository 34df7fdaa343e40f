"""GPU parity of the stand-alone MMA functions (reference names and signatures) and of the
incremental decoding step, against golden vectors from the unmodified reference and the oracle."""
import numpy as np
import pytest
import torch

from oracle import mma as omma
from tests.golden_io import GOLDEN_DIR, load, opt
from tests.test_mma_train_gpu import assert_parity

pytestmark = pytest.mark.gpu

TRAIN = load("mma_train.npz")
STEP = load("mma_step.npz")
DEV = "cuda"


@pytest.mark.parametrize("name", list(TRAIN))
def test_three_functions_called_separately_match_golden(name):
    """expected_alignment_from_p_choose -> mass_preservation -> expected_soft_attention, each its
    own autograd node, exactly how the reference module chains them (:318-347)."""
    from simulst_b200.utils import monotonic_attention as ma
    c = TRAIN[name]
    n, t, s, masked, chunk, soft, mp = [int(v) for v in c.cfg]
    p = c.p.to(DEV).requires_grad_()
    se = c.soft_energy.to(DEV).requires_grad_()
    mask = opt(c.mask)
    mask = mask.to(DEV) if mask is not None else None
    alpha = ma.expected_alignment_from_p_choose(p.float(), mask, eps=1e-6)
    if mp:
        alpha = ma.mass_preservation(alpha, mask)
    if soft:
        beta = ma.expected_soft_attention(alpha, se, padding_mask=mask, chunk_size=chunk or None, eps=1e-6)
        loss = (alpha * c.g_alpha.to(DEV)).sum() + (beta * c.g_beta.to(DEV)).sum()
    else:
        beta = alpha
        loss = (alpha * c.g_alpha.to(DEV)).sum()
    loss.backward()
    assert_parity(alpha.detach().cpu(), c.alpha, "alpha")
    assert_parity(beta.detach().cpu(), c.beta, "beta")
    assert_parity(p.grad.cpu(), c.grad_p, "grad_p")
    if soft:
        assert_parity(se.grad.cpu(), c.grad_soft_energy, "grad_soft_energy")


def test_mass_preservation_is_in_place_without_mask():
    from simulst_b200.utils import monotonic_attention as ma
    a = torch.rand(2, 3, 10, device=DEV) * 0.05
    ref = omma.mass_preservation(a.cpu().clone())
    out = ma.mass_preservation(a)
    assert out.data_ptr() == a.data_ptr()
    torch.testing.assert_close(a.cpu(), ref, rtol=1e-6, atol=1e-7)


def test_moving_sum_docstring_example_and_random():
    from simulst_b200.utils.functions import moving_sum
    z = np.load(GOLDEN_DIR + "/moving_sum.npz")
    x = torch.from_numpy(z["x"]).to(DEV)
    assert torch.equal(moving_sum(x, 3, 1)[0].cpu(), torch.from_numpy(z["doc_s3e1"]))
    assert torch.equal(moving_sum(x, 1, 3)[0].cpu(), torch.from_numpy(z["doc_s1e3"]))
    g = torch.Generator().manual_seed(3)
    y = torch.randn(3, 4, 77, generator=g)
    for a, b in ((5, 1), (1, 5), (4, 3)):
        torch.testing.assert_close(moving_sum(y.to(DEV), a, b).cpu(), omma.moving_sum(y, a, b),
                                   rtol=1e-6, atol=1e-6)


def test_exclusive_and_safe_cumprod():
    from simulst_b200.utils.functions import exclusive_cumprod, safe_cumprod
    g = torch.Generator().manual_seed(4)
    x = torch.rand(3, 5, 130, generator=g)
    got = exclusive_cumprod(x.to(DEV), dim=2, eps=1e-6).cpu()
    ref = omma.exclusive_cumprod(x, dim=2, eps=1e-6)
    assert_parity(got, ref, "exclusive_cumprod")
    assert abs(float(got[0, 0, 0]) - 1.00000095) < 1e-7          # first element is 1 + eps, not 1
    assert_parity(safe_cumprod(x.to(DEV), dim=2, eps=1e-6).cpu(), omma.safe_cumprod(x, 2, 1e-6), "safe")
    assert_parity(exclusive_cumprod(x.to(DEV), dim=1, eps=1e-6).cpu(),
                  omma.exclusive_cumprod(x, dim=1, eps=1e-6), "dim1")
    with pytest.raises(RuntimeError, match="non-negative"):
        safe_cumprod(-torch.ones(1, 1, 4, device=DEV), dim=2)


def test_learnable_p_choose():
    from simulst_b200.utils.p_choose_strategy import learnable_p_choose
    g = torch.Generator().manual_seed(5)
    e = torch.randn(2, 3, 50, generator=g)
    got = learnable_p_choose(e.to(DEV), training=False).cpu()
    torch.testing.assert_close(got, torch.sigmoid(e), rtol=1e-6, atol=1e-7)
    torch.manual_seed(11)
    got = learnable_p_choose(e.to(DEV), 0.5, 2.0, training=True).cpu()
    torch.manual_seed(11)
    noise = torch.randn_like(e.to(DEV)).cpu() * 2.0 + 0.5
    torch.testing.assert_close(got, torch.sigmoid(e + noise), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(STEP))
def test_incremental_step_matches_golden(name):
    """head_step / head_read / one-hot alpha bit-exact, beta rtol 1e-5."""
    from simulst_b200 import ops
    c = STEP[name]
    soft, mp, bsz, heads, s, steps, masked = [int(v) for v in c.cfg]
    n = bsz * heads
    lens = c.src_lengths.to(DEV) if masked else None
    head_step = torch.zeros(n, dtype=torch.long, device=DEV)
    for st in range(steps):
        se = c.soft_energy[st].to(DEV) if soft else None
        head_read, alpha, beta = ops.mma_step(c.p[st].to(DEV), head_step, se, lens, bool(mp))
        assert torch.equal(head_step.cpu(), c.head_step[st]), st
        assert torch.equal(head_read.cpu(), c.head_read[st]), st
        assert torch.equal(alpha.cpu(), c.alpha[st]), st
        if soft:
            torch.testing.assert_close(beta.cpu(), c.beta[st], rtol=1e-5, atol=1e-7)


def test_incremental_step_256_utterances():
    """BASELINE config 4 shape: 256 concurrent utterances x 4 heads, 32 consecutive steps."""
    from simulst_b200 import ops
    bsz, heads, s = 256, 4, 256
    n = bsz * heads
    g = torch.Generator().manual_seed(3000)
    lens = torch.randint(s // 2, s + 1, (bsz,), generator=g).repeat_interleave(heads)
    mask = torch.arange(s)[None, :] >= lens[:, None]
    hs_o = torch.zeros(n, dtype=torch.long)
    hs_k = torch.zeros(n, dtype=torch.long, device=DEV)
    for st in range(32):
        p = torch.sigmoid(torch.randn(n, s, generator=g) - 2.0).masked_fill(mask, 0.0)
        se = torch.randn(n, s, generator=g).masked_fill(mask, -1e8)
        hs_o, hr_o, a_o, b_o = omma.mma_process_infer(p, hs_o, se.unsqueeze(1), mask, True)
        hr_k, a_k, b_k = ops.mma_step(p.to(DEV), hs_k, se.to(DEV), lens.to(DEV), True)
        assert torch.equal(hs_k.cpu(), hs_o) and torch.equal(hr_k.cpu(), hr_o)
        assert torch.equal(a_k.cpu(), a_o)
        torch.testing.assert_close(b_k.cpu(), b_o.squeeze(1), rtol=1e-5, atol=1e-7)


def test_module_mixin_train_and_infer_against_oracle():
    """The mixin drives the kernels through the reference module's own attribute protocol."""
    from simulst_b200.modules.monotonic_multihead_attention import B200MonotonicAttentionMixin

    class Host(B200MonotonicAttentionMixin):
        num_heads, eps, mass_preservation, soft_attention, chunk_size = 2, 1e-6, True, True, None

        def __init__(self, p, e):
            self._p, self._e, self.state = p, e, {}

        def p_choose(self, q, k, m, inc=None):
            return self._p

        def energy_from_qk(self, q, k, kind, key_padding_mask=None, bias=0):
            return self._e

        def _get_monotonic_buffer(self, inc):
            return self.state

        def _set_monotonic_buffer(self, inc, buf):
            self.state = buf

    g = torch.Generator().manual_seed(8)
    bsz, t, s = 3, 5, 48
    p = torch.sigmoid(torch.randn(bsz * 2, t, s, generator=g) - 2)
    e = torch.randn(bsz * 2, t, s, generator=g)
    host = Host(p.to(DEV), e.to(DEV))
    q = torch.zeros(t, bsz, 8, device=DEV)
    k = torch.zeros(s, bsz, 8, device=DEV)
    _, alpha, beta, _ = host.monotonic_attention_process_train(q, k, None)
    a_o, b_o = omma.mma_process_train(p, e, None, 1e-6, True, None)
    assert_parity(alpha.cpu(), a_o, "alpha")
    assert_parity(beta.cpu(), b_o, "beta")
    assert host.expected_delays is None
    host.with_expected_delays = True            # opt-in latency-loss epilogue (mma_criterion.py:146-157)
    _, alpha_d, _, _ = host.monotonic_attention_process_train(q, k, None)
    assert torch.equal(alpha_d, alpha)
    assert_parity(host.expected_delays.cpu(), omma.expected_delays(a_o), "expected_delays")
    host2 = Host(p[:, :1].to(DEV), e[:, :1].to(DEV))
    _, alpha1, beta1 = host2.monotonic_attention_process_infer(q[:1], k, None, {})
    ns, hr, a1, b1 = omma.mma_process_infer(p[:, 0], torch.zeros(bsz * 2, dtype=torch.long), e[:, :1], None, True)
    assert torch.equal(host2.state["head_step"].cpu().view(-1), ns)
    assert torch.equal(host2.state["head_read"].cpu().view(-1), hr)
    assert torch.equal(alpha1.cpu(), a1)
    torch.testing.assert_close(beta1.cpu(), b1, rtol=1e-5, atol=1e-7)


def test_moving_sum_and_cumprod_are_differentiable():
    """the reference's moving_sum (conv1d) and exclusive/safe_cumprod (log-cumsum-exp)
    are differentiable; the mirrors must not drop the graph."""
    from simulst_b200.utils.functions import exclusive_cumprod, moving_sum, safe_cumprod
    g = torch.Generator().manual_seed(6)
    x = torch.rand(2, 3, 41, generator=g) * 0.9 + 0.05
    w = torch.randn(2, 3, 41, generator=g)
    for fn_k, fn_o in ((lambda t: moving_sum(t, 4, 2), lambda t: omma.moving_sum(t, 4, 2)),
                       (lambda t: moving_sum(t, 1, 5), lambda t: omma.moving_sum(t, 1, 5)),
                       (lambda t: exclusive_cumprod(t, dim=2, eps=1e-6), lambda t: omma.exclusive_cumprod(t, 2, 1e-6)),
                       (lambda t: safe_cumprod(t, dim=2, eps=1e-6), lambda t: omma.safe_cumprod(t, 2, 1e-6)),
                       (lambda t: exclusive_cumprod(t, dim=1, eps=1e-6), lambda t: omma.exclusive_cumprod(t, 1, 1e-6))):
        x_k = x.to(DEV).requires_grad_()
        x_o = x.clone().requires_grad_()
        y_k = fn_k(x_k)
        assert y_k.grad_fn is not None
        (y_k * w.to(DEV)).sum().backward()
        (fn_o(x_o) * w).sum().backward()
        assert_parity(x_k.grad, x_o.grad, "grad", atol=2e-6)


def test_cumprod_check_ignores_stale_status_bits():
    """safe_cumprod's negative-input check uses a status word of its own."""
    import simulst_b200
    from simulst_b200 import ops
    from simulst_b200.utils.functions import safe_cumprod
    bad = torch.rand(1, 2, 16, device=DEV)
    bad[0, 0, 3] = 1.5
    ops.mma_train(bad, None, None, mass_preservation=False)        # leaves ST_RANGE on the device word
    out = safe_cumprod(torch.rand(1, 1, 8, device=DEV), dim=2)     # must not raise for it
    assert out.shape == (1, 1, 8)
    with pytest.raises(AssertionError, match="Incorrect values"):
        simulst_b200.check_status(torch.device(DEV))
