"""SSNT lattice loss: oracle pinned by golden vectors from the unmodified
reference (tests/golden/ssnt.npz) and by the reference's own brute-force checker
(ssnt_loss/test.py:19-80, restated in oracle.ssnt); kernels against both."""
import pytest
import torch

from oracle import ssnt as ossnt
from tests.golden_io import load
from tests.parity import assert_parity

SSNT = load("ssnt.npz")


def _case(c):
    n, t, s, v, use_logits = [int(x) for x in c.cfg]
    return n, t, s, v, bool(use_logits), float(c.fastemit), str(c.reduction)


@pytest.mark.parametrize("name", list(SSNT))
def test_ssnt_oracle_matches_reference_golden(name):
    c = SSNT[name]
    n, t, s, v, use_logits, lam, red = _case(c)
    logits = c.logits.clone().requires_grad_()
    emit = c.emit.clone().requires_grad_()
    kw = {"emit_logits": emit} if use_logits else {"emit_probs": emit}
    loss, lattice, log_p = ossnt.ssnt_loss(logits.log_softmax(-1), c.targets, c.source_lengths, c.target_lengths,
                                           reduction=red, fastemit_lambda=lam, **kw)
    w = c.w if isinstance(c.w, torch.Tensor) else torch.tensor(c.w)
    (loss * w).sum().backward()
    tight = dict(rtol=1e-6, atol=1e-5)
    torch.testing.assert_close(loss, torch.as_tensor(c.loss), **tight)
    torch.testing.assert_close(lattice, c.lattice, **tight)
    torch.testing.assert_close(log_p, c.log_p_choose, **tight)
    torch.testing.assert_close(emit.grad, c.grad_emit, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits.grad, c.grad_logits, rtol=1e-5, atol=1e-6)
    keep = torch.arange(t)[None, :] < c.target_lengths[:, None]
    kw_m = {k: x.detach()[keep] for k, x in kw.items()}
    loss_m, lat_m, _ = ossnt.ssnt_loss_mem(logits.detach().log_softmax(-1)[keep], c.targets[keep], c.source_lengths,
                                           c.target_lengths, reduction=red, fastemit_lambda=lam, **kw_m)
    torch.testing.assert_close(loss_m, torch.as_tensor(c.loss_mem), **tight)
    torch.testing.assert_close(lat_m, c.lattice_mem, **tight)


def test_ssnt_oracle_matches_bruteforce_checker():
    """The reference's own acceptance criterion: lattice and loss within 1e-3 of the triple loop."""
    g = torch.Generator().manual_seed(71)
    n, t, s, v = 2, 4, 9, 5
    lp = torch.rand(n, t, s, v, generator=g).log_softmax(-1)
    emit = torch.rand(n, t, s, generator=g)
    targets = torch.randint(0, v, (n, t), generator=g)
    src, tgt = torch.tensor([9, 6]), torch.tensor([4, 3])
    loss, lattice, _ = ossnt.ssnt_loss(lp, targets, src, tgt, emit_logits=emit)
    loss_b, lat_b = ossnt.ssnt_lattice_bruteforce(lp, targets, src, tgt, emit.sigmoid())
    torch.testing.assert_close(loss, loss_b.float(), rtol=1e-3, atol=1e-3)
    for b in range(n):      # the checker knows no padding beyond source_len / target_len
        torch.testing.assert_close(lattice[b, :, :src[b]], lat_b[b, :, :src[b]].float(), rtol=1e-3, atol=1e-3)


def _kernel_run(c, flat, dtype=torch.float32):
    from simulst_b200.criterion import ssnt_loss as kssnt
    n, t, s, v, use_logits, lam, red = _case(c)
    dev = "cuda"
    logits = c.logits.to(dev).requires_grad_()
    emit = c.emit.to(dev).requires_grad_()
    lp = logits.log_softmax(-1)
    if flat:
        keep = (torch.arange(t)[None, :] < c.target_lengths[:, None]).to(dev)
        kw = {"emit_logits" if use_logits else "emit_probs": emit[keep]}
        out = kssnt.ssnt_loss_mem(lp[keep], c.targets.to(dev)[keep], c.source_lengths.to(dev),
                                  c.target_lengths.to(dev), reduction=red, fastemit_lambda=lam, **kw)
    else:
        kw = {"emit_logits" if use_logits else "emit_probs": emit}
        out = kssnt.ssnt_loss(lp, c.targets.to(dev), c.source_lengths.to(dev), c.target_lengths.to(dev),
                              reduction=red, fastemit_lambda=lam, **kw)
    return out, logits, emit


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SSNT))
@pytest.mark.parametrize("flat", [False, True], ids=["padded", "flat"])
def test_ssnt_kernel_matches_reference_golden(name, flat):
    import simulst_b200
    c = SSNT[name]
    n, t, s, v, use_logits, lam, red = _case(c)
    (loss, lattice, log_p), logits, emit = _kernel_run(c, flat)
    w = c.w if isinstance(c.w, torch.Tensor) else torch.tensor(c.w)
    (loss * w.to(loss.device)).sum().backward()
    simulst_b200.check_status(loss.device)
    # log-space values reach |neg_inf| = 1e4: rtol 1e-5 of the value, floor 1e-6 x max|lattice|
    if flat:
        assert_parity(loss.view(-1), torch.as_tensor(c.loss_mem).view(-1), f"{name} loss (flat)")
        assert_parity(lattice, c.lattice_mem, f"{name} lattice (flat)")
    else:
        assert_parity(loss.view(-1), torch.as_tensor(c.loss).view(-1), f"{name} loss")
        assert_parity(lattice, c.lattice, f"{name} lattice")
        assert_parity(log_p, c.log_p_choose, f"{name} log_p_choose")
    # gradients are probabilities (<= |w|): absolute floor 2e-6 x max|w|
    floor = 2e-6 * float(w.abs().max())
    assert_parity(emit.grad, c.grad_emit, f"{name} grad_emit", extra_atol=floor)
    assert_parity(logits.grad, c.grad_logits, f"{name} grad_logits", extra_atol=floor)


@pytest.mark.gpu
def test_ssnt_kernel_lattice_and_log_p_gradients():
    """Gradients that enter through the lattice and through log_p_choose (the criterion's offline
    loss reads lprobs_emit, ssnt_criterion.py:160-183), mixed signs, against the oracle's autograd."""
    from simulst_b200.criterion import ssnt_loss as kssnt
    g = torch.Generator().manual_seed(72)
    n, t, s, v = 3, 6, 70, 13
    logits = torch.randn(n, t, s, v, generator=g)
    emit = torch.randn(n, t, s, generator=g) - 1
    targets = torch.randint(0, v, (n, t), generator=g)
    src, tgt = torch.tensor([70, 41, 55]), torch.tensor([6, 4, 5])
    w_lat = torch.randn(n, t, s, generator=g) * 0.1
    w_lp = torch.randn(n, t, s, generator=g) * 0.1
    lo, eo = logits.clone().requires_grad_(), emit.clone().requires_grad_()
    loss, lat, lpc = ossnt.ssnt_loss(lo.log_softmax(-1), targets, src, tgt, emit_logits=eo, reduction="sum")
    (loss + (lat * w_lat).sum() + (lpc * w_lp).sum()).backward()
    lk, ek = logits.cuda().requires_grad_(), emit.cuda().requires_grad_()
    loss_k, lat_k, lpc_k = kssnt.ssnt_loss(lk.log_softmax(-1), targets.cuda(), src.cuda(), tgt.cuda(),
                                           emit_logits=ek, reduction="sum")
    (loss_k + (lat_k * w_lat.cuda()).sum() + (lpc_k * w_lp.cuda()).sum()).backward()
    assert_parity(loss_k.view(1), loss.detach().view(1), "loss")
    assert_parity(lat_k, lat.detach(), "lattice")
    assert_parity(ek.grad, eo.grad, "grad_emit", extra_atol=2e-6)
    assert_parity(lk.grad, lo.grad, "grad_logits", extra_atol=2e-6)


@pytest.mark.gpu
def test_ssnt_kernel_bf16_inputs_and_range_check():
    import simulst_b200
    from simulst_b200.criterion import ssnt_loss as kssnt
    g = torch.Generator().manual_seed(73)
    n, t, s, v = 2, 5, 300, 17
    lp = torch.randn(n, t, s, v, generator=g).log_softmax(-1).to(torch.bfloat16)
    emit = (torch.randn(n, t, s, generator=g) - 1).to(torch.bfloat16)
    targets = torch.randint(0, v, (n, t), generator=g)
    src, tgt = torch.tensor([300, 211]), torch.tensor([5, 3])
    loss_o, lat_o, _ = ossnt.ssnt_loss(lp.float(), targets, src, tgt, emit_logits=emit.float())
    loss_k, lat_k, _ = kssnt.ssnt_loss(lp.cuda(), targets.cuda(), src.cuda(), tgt.cuda(), emit_logits=emit.cuda())
    assert lat_k.dtype == torch.float32
    assert_parity(loss_k, loss_o, "bf16 loss", rtol=2e-5)
    assert_parity(lat_k, lat_o, "bf16 lattice", rtol=2e-5)
    simulst_b200.check_status(torch.device("cuda"))
    bad = lp.float().cuda().clone()
    bad[0, 0, 0, 0] = 0.5                   # a "log-probability" > 0
    kssnt.ssnt_loss(bad, targets.cuda(), src.cuda(), tgt.cuda(), emit_logits=emit.float().cuda())
    with pytest.raises(AssertionError, match="Incorrect values"):
        simulst_b200.check_status(torch.device("cuda"))
