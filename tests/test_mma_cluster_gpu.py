"""Forward of long rows in small batches on thread-block clusters (csrc/mma_fwd_cluster.cuh): the cluster kernel
performs the single-CTA kernel's arithmetic in the same order, so the two must agree BIT FOR BIT (alpha, beta,
expected delays, the mass-preservation side values through the backward); one shape is also held against the
CPU oracle under the shared parity gate."""
import pytest
import torch

from oracle import mma as omma
from tests.parity import assert_parity
from tests.test_mma_train_gpu import _seeded

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fwd_bwd(p, se, mp, ga, gb, dtype, delays=False):
    from simulst_b200 import ops
    p_d = p.to(DEV, dtype).requires_grad_()
    se_d = se.to(DEV, dtype).requires_grad_() if se is not None else None
    if delays:
        alpha, beta, d = ops.mma_train_with_delays(p_d, se_d, None, eps=1e-6, mass_preservation=mp)
    else:
        alpha, beta = ops.mma_train(p_d, se_d, None, eps=1e-6, mass_preservation=mp)
        d = None
    loss = (alpha * ga.to(DEV)).sum()
    if se is not None:
        loss = loss + (beta * gb.to(DEV)).sum()
    if d is not None:
        loss = loss + d.sum()
    loss.backward()
    torch.cuda.synchronize()
    return alpha.detach(), beta.detach(), d, p_d.grad, (se_d.grad if se is not None else None)


@pytest.fixture
def cluster_mode():
    from simulst_b200 import _lib
    lib = _lib.load()

    def set_mode(m):
        assert lib.simulst_mma_set_cluster(m) == 0
    yield set_mode
    lib.simulst_mma_set_cluster(1)


SHAPES = [
    # n, T, S                      cluster shape
    (3, 9, 1280),                # 2 x 96 threads, ragged last slice
    (2, 7, 2048),                # 2 x 128
    (3, 5, 2056),                # 4 x 96, the last CTA holds no column
    (2, 6, 3000),                # 4 x 96
    (2, 40, 4096),               # 4 x 128, deep enough to reuse every ring slot and both exchange phases many times
    (1, 5, 6000),                # 8 x 96
    (1, 4, 8192),                # 8 x 128
    (75, 3, 2048),               # more rows than the automatic mode takes: forced
]


@pytest.mark.parametrize("n,t,s", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("soft,mp,delays", [(True, True, False), (True, False, True), (False, True, True), (False, False, False)])
def test_cluster_forward_is_bit_identical(n, t, s, dtype, soft, mp, delays, cluster_mode):
    p, se, _, ga, gb = _seeded(n, t, s, seed=500 + s + t)
    se_in = se if soft else None
    cluster_mode(0)
    ref = _fwd_bwd(p, se_in, mp, ga, gb, dtype, delays)
    cluster_mode(2)
    got = _fwd_bwd(p, se_in, mp, ga, gb, dtype, delays)
    for name, a, b in zip(("alpha", "beta", "delays", "grad_p", "grad_energy"), got, ref):
        if a is None:
            assert b is None
            continue
        if s <= 4096:
            assert torch.equal(a, b), f"{name}: cluster kernel differs from the single-CTA kernel (max {float((a.float() - b.float()).abs().max()):.3e})"
        else:
            # the single-CTA kernel runs such rows with 12 or 16 elements per thread: another summation order
            # (forward outputs only: the backward is the same kernel in both runs)
            scale = max(float(b.detach().float().abs().max()), 1e-30)
            if name in ("alpha", "beta"):
                torch.testing.assert_close(a.float(), b.float(), rtol=1e-5, atol=2e-6 * scale)
            elif name == "delays":
                torch.testing.assert_close(a.float(), b.float(), rtol=1e-4, atol=1e-5 * scale)


def test_cluster_forward_matches_oracle(cluster_mode):
    n, t, s = 2, 24, 4096
    p, se, _, ga, gb = _seeded(n, t, s, seed=4096)
    p, se = p.bfloat16(), se.bfloat16()
    cluster_mode(1)             # automatic: 2 rows <= 74
    alpha, beta, _, gp, ge = _fwd_bwd(p, se, True, ga, gb, torch.bfloat16)
    outs = {}
    for dt in (torch.float32, torch.float64):
        p_o = p.to(dt).requires_grad_()
        se_o = se.to(dt).requires_grad_()
        a_o, b_o = omma.mma_process_train(p_o, se_o, None, 1e-6, True, None, compute_dtype=dt)
        ((a_o * ga).sum() + (b_o * gb).sum()).backward()
        outs[dt] = (a_o.detach(), b_o.detach(), p_o.grad, se_o.grad)
    a_o, b_o, gp_o, ge_o = outs[torch.float32]
    a64, b64, gp64, ge64 = outs[torch.float64]
    assert_parity(alpha, a_o, "cluster alpha", a64)
    assert_parity(beta, b_o, "cluster beta", b64)
    floor = 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(ga.abs().max()), float(gb.abs().max()))
    assert_parity(gp.float(), gp_o, "cluster grad_p", gp64, rtol=2.0 ** -8, extra_atol=floor)
    assert_parity(ge.float(), ge_o, "cluster grad_energy", ge64, rtol=2.0 ** -8, extra_atol=floor)


def test_cluster_forward_status_word(cluster_mode):
    """prob_check through the status word on the cluster path."""
    import simulst_b200
    from simulst_b200 import ops
    n, t, s = 2, 4, 4096
    p, se, _, _, _ = _seeded(n, t, s, seed=1)
    cluster_mode(2)
    simulst_b200.check_status()
    ops.mma_train(p.to(DEV), se.to(DEV), None)
    simulst_b200.check_status()
    bad = p.clone()
    bad[1, 2, 3000] = float("nan")
    ops.mma_train(bad.to(DEV), se.to(DEV), None)
    with pytest.raises(AssertionError):
        simulst_b200.check_status()
