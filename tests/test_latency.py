"""Latency loss: oracle and kernels against tests/golden/latency.npz, which
is produced by the reference's own MMACriterion.compute_latency_loss source
(criterion/mma_criterion.py:138-207) with SimulEval's DAL restated (oracle/latency.py)."""
import pytest
import torch

from oracle import latency as olat
from oracle import mma as omma
from tests.golden_io import load
from tests.parity import assert_parity

LAT = load("latency.npz")


def _unpack(c):
    bsz, layers, heads, t, s = [int(v) for v in c.cfg]
    avg_w, var_w = [float(v) for v in c.weights]
    cfg = olat.criterion_stub(avg_w, var_w, str(c.gather), 1, 10.0)
    return bsz, layers, heads, t, s, cfg


@pytest.mark.parametrize("name", list(LAT))
def test_oracle_latency_loss_matches_reference_golden(name):
    c = LAT[name]
    bsz, layers, heads, t, s, cfg = _unpack(c)
    mask_h = torch.repeat_interleave(c.enc_mask, heads, 0)
    ps, alphas = [], []
    for l in range(layers):
        p = c.p[l].clone().requires_grad_()
        a, _ = omma.mma_process_train(p, None, mask_h, 1e-6, True, None)
        ps.append(p)
        alphas.append(a.view(bsz, heads, t, s))
    loss, latency, var, delays = olat.mma_latency_loss(alphas, c.target, c.src_lengths, c.enc_mask, cfg)
    loss.backward()
    tight = dict(rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(delays, c.expected_delays, **tight)
    torch.testing.assert_close(loss, torch.tensor(c.latency_loss), **tight)
    torch.testing.assert_close(latency, torch.tensor(c.expected_latency), **tight)
    torch.testing.assert_close(var, torch.tensor(c.delays_var), **tight)
    for l in range(layers):
        torch.testing.assert_close(ps[l].grad, c.grad_p[l], rtol=1e-5, atol=1e-6)


def test_dal_closed_form():
    """DAL of a constant-lag policy: g(i) = k + i/gamma -> g' = g, DAL = k (paper eq. 20-21)."""
    t, src = 6, 12
    gamma = t / src
    d = (3.0 + torch.arange(t) / gamma).view(1, t)
    out = olat.differentiable_average_lagging(d, torch.tensor([src]))
    torch.testing.assert_close(out, torch.tensor([[3.0]]))
    # a late first write drags every later step: g' = g(0) + i/gamma
    d2 = torch.tensor([[7.0, 1.0, 2.0, 3.0, 4.0, 5.0]])
    out2 = olat.differentiable_average_lagging(d2, torch.tensor([src]))
    torch.testing.assert_close(out2, torch.tensor([[7.0]]))


@pytest.mark.gpu
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("with_ref", [False, True])
def test_dal_kernel_matches_oracle(masked, with_ref):
    from simulst_b200 import ops
    g = torch.Generator().manual_seed(61)
    n, t = 37, 53
    d = (torch.rand(n, t, generator=g) * 40).cumsum(1) * 0.2
    d[3] = d[3].flip(0)                    # non-monotone delays: the carried term wins
    src = torch.randint(30, 200, (n,), generator=g)
    ref = torch.randint(5, t + 1, (n,), generator=g) if with_ref else None
    mask = None
    if masked:
        tl = torch.randint(1, t + 1, (n,), generator=g)
        mask = torch.arange(t)[None, :] >= tl[:, None]
    gout = torch.randn(n, 1, generator=g)
    d_o = d.clone().requires_grad_()
    out_o = olat.differentiable_average_lagging(d_o, src, ref, mask)
    (out_o * gout).sum().backward()
    d_k = d.cuda().requires_grad_()
    out_k = ops.differentiable_average_lagging(d_k, src.cuda(), ref.cuda() if with_ref else None,
                                               mask.cuda() if masked else None)
    (out_k * gout.cuda()).sum().backward()
    assert_parity(out_k, out_o.detach(), "dal")
    assert_parity(d_k.grad, d_o.grad, "grad_delays")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(LAT))
@pytest.mark.parametrize("fused_delays", [True, False])
def test_kernel_latency_loss_matches_reference_golden(name, fused_delays):
    """alignment kernel (+ expected-delay epilogue) -> simulst_dal -> gather / variance, through
    the criterion-shaped mirror, against what the reference's method returned; gradients all the
    way back to p_choose."""
    from simulst_b200 import ops
    from simulst_b200.criterion.mma_latency import compute_latency_loss
    c = LAT[name]
    bsz, layers, heads, t, s, cfg = _unpack(c)
    dev = "cuda"
    mask_h = torch.repeat_interleave(c.enc_mask, heads, 0).to(dev)
    ps, alphas, delays = [], [], []
    for l in range(layers):
        p = c.p[l].to(dev).requires_grad_()
        a, _, d = ops.mma_train_with_delays(p, None, mask_h, eps=1e-6, mass_preservation=True)
        ps.append(p)
        alphas.append(a.view(bsz, heads, t, s))
        delays.append(d)
    sample = {"target": c.target.to(dev), "net_input": {"src_lengths": c.src_lengths.to(dev)}}
    net_output = (None, {"attn_list": [{"alpha": a} for a in alphas], "encoder_padding_mask": [c.enc_mask.to(dev)]})
    loss, latency, var = compute_latency_loss(cfg, None, sample, net_output, delays if fused_delays else None)
    loss.backward()
    assert_parity(torch.cat([d.view(bsz, heads, t) for d in delays], 1).reshape(-1, t), c.expected_delays, "delays")
    assert_parity(loss.view(1), torch.tensor([c.latency_loss]), "latency_loss")
    assert_parity(latency.view(1), torch.tensor([c.expected_latency]), "expected_latency")
    assert_parity(var.view(1), torch.tensor([c.delays_var]), "delays_var",
                  extra_atol=1e-5 * float(c.expected_delays.abs().max()))
    for l in range(layers):
        assert_parity(ps[l].grad, c.grad_p[l], f"grad_p layer {l}", extra_atol=2e-6 * s)
