"""Expected-delay epilogue of the fused MMA training op: the kernel's
`sum_j (j+1) * alpha'_ij` against the reference expression
(codebase/criterion/mma_criterion.py:146-157, restated in oracle/mma.py::expected_delays) evaluated
on the oracle's alpha, forward and backward, through the Python mirror / C ABI, for every kernel
family.  Tolerance: 1e-5 relative plus an absolute floor tied to the tensor's scale (see
tests/test_mma_train_gpu.py::assert_parity); the delays sum up to S terms of size <= S, so their
floor is 1e-5 of the largest delay."""
import pytest
import torch

from oracle import mma as omma
from tests.test_mma_train_gpu import _seeded, assert_parity, kernel_family  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

CASES = [
    # n, t, s, masked, chunk, soft, mass preservation
    (8, 32, 256, False, 0, True, True),        # BASELINE config 1 rows
    (4, 9, 1024, False, 0, True, True),        # dense fast-path backward
    (4, 9, 512, True, 0, True, True),          # right padding: residual ADDED at src_len-1
    (3, 5, 264, False, 0, False, False),       # hard attention, no mass preservation
    (3, 5, 1000, False, 0, False, True),       # hard attention, ragged row
    (2, 3, 4096, False, 0, True, True),        # 16 warps
    (2, 3, 6000, True, 0, True, True),         # long-form, masked (generic kernels)
    (3, 6, 300, False, 7, True, True),         # chunkwise (generic kernels)
]


def _run(p, se, mask, mp, chunk, soft, w_d, g_beta, g_alpha=None):
    from simulst_b200 import ops
    dev = torch.device("cuda")
    p_d = p.detach().to(dev).requires_grad_()
    se_d = se.detach().to(dev).requires_grad_() if soft else None
    m_d = mask.to(dev) if mask is not None else None
    alpha, beta, delays = ops.mma_train_with_delays(p_d, se_d, m_d, eps=1e-6, mass_preservation=mp,
                                                    chunk_size=chunk or None)
    loss = (delays * w_d.to(dev)).sum()
    if soft:
        loss = loss + (beta * g_beta.to(dev)).sum()
    if g_alpha is not None:
        loss = loss + (alpha * g_alpha.to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    return (alpha.detach().cpu(), delays.detach().cpu(), p_d.grad.cpu(), se_d.grad.cpu() if soft else None)


def _oracle(p, se, mask, mp, chunk, soft, w_d, g_beta, g_alpha=None, dtype=torch.float32):
    p_o = p.detach().clone().to(dtype).requires_grad_()
    se_o = se.detach().clone().to(dtype).requires_grad_() if soft else None
    kw = {} if dtype == torch.float32 else {"compute_dtype": dtype}
    a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, mp, chunk or None, **kw)
    d_o = omma.expected_delays(a_o)
    loss = (d_o * w_d).sum()
    if soft:
        loss = loss + (b_o * g_beta).sum()
    if g_alpha is not None:
        loss = loss + (a_o * g_alpha).sum()
    loss.backward()
    return a_o.detach(), d_o.detach(), p_o.grad, se_o.grad if soft else None


@pytest.mark.parametrize("n,t,s,masked,chunk,soft,mp", CASES)
@pytest.mark.parametrize("with_alpha_grad", [False, True])
def test_expected_delays_match_reference_expression(n, t, s, masked, chunk, soft, mp, with_alpha_grad, kernel_family):
    p, se, mask, ga, gb = _seeded(n, t, s, seed=300 + s + t, masked=masked)
    g = torch.Generator().manual_seed(7)
    w_d = torch.randn(n, t, generator=g) / s            # latency-loss sized weights
    ga = ga if with_alpha_grad else None
    alpha, delays, gp, ge = _run(p, se, mask, mp, chunk, soft, w_d, gb, ga)
    a_o, d_o, gp_o, ge_o = _oracle(p, se, mask, mp, chunk, soft, w_d, gb, ga)
    a64, d64, gp64, ge64 = _oracle(p, se, mask, mp, chunk, soft, w_d, gb, ga, dtype=torch.float64)
    assert_parity(alpha, a_o, "alpha", a64)
    assert_parity(delays, d_o, "expected_delays", d64)
    # the delay term injects gd_i*(j+1) (up to max|w_d|*S) into dL/dalpha; the suffix sums that
    # follow cancel at that scale, so the absolute floor is tied to it, not to the (smaller) result
    inj = 1e-6 * float(w_d.abs().max()) * s
    assert_parity(gp, gp_o, "grad_p", gp64, extra_atol=inj)
    if soft:
        assert_parity(ge, ge_o, "grad_soft_energy", ge64, extra_atol=inj)


def test_delays_equal_weighted_row_sum_of_the_returned_alpha():
    """Size-independent property at the training shape's row length: the epilogue equals the
    weighted row sum of the alpha the same launch returned (fp64 sum of the fp32 alpha)."""
    from simulst_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    p = torch.sigmoid(torch.randn(16, 128, 1024, generator=g) - 2.0).to(dev, torch.bfloat16)
    e = torch.randn(16, 128, 1024, generator=g).to(dev, torch.bfloat16)
    alpha, beta, delays = ops.mma_train_with_delays(p, e, None)
    steps = torch.arange(1, 1025, device=dev, dtype=torch.float64)
    ref = (alpha.double() * steps).sum(-1)
    torch.testing.assert_close(delays.double(), ref, rtol=1e-5, atol=1e-5 * float(ref.max()))
