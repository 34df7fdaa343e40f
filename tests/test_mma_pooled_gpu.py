"""SURVEY 8f #2 -- pooled p_choose producer (fixed pre-decision): ``simulst_mma_train_{fwd,bwd}_pooled``
through ``ops.mma_train_pooled`` against

* the golden vectors the reference's own ``*_fixed_pre_decision`` classes produced
  (tests/golden/fixed_predecision.npz, make_golden_r2.py::gen_fixed_predecision),
* the oracle (``oracle.mma.mma_process_train_pooled``) on seeded inputs at the training shape,
* the dense kernels fed the expanded row (bit-identical: same arithmetic, only the load differs),
* the real wrapper class driven through the reference's ``forward()``."""
import copy

import pytest
import torch

from oracle import mma as omma
from oracle import ref_loader
from tests.golden_io import load, opt
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"
FIXED = load("fixed_predecision.npz")


def _run(pp, s, ratio, se, mask, mp, ga, gb, dtype=torch.float32, right_padding=False, want_dense=True,
         chunk=None):
    from simulst_b200 import ops
    ppd = pp.to(DEV, dtype).requires_grad_()
    sed = se.to(DEV, dtype).requires_grad_() if se is not None else None
    md = mask.to(DEV) if mask is not None else None
    dense, alpha, beta, _ = ops.mma_train_pooled(ppd, s, ratio, sed, md, eps=1e-6, mass_preservation=mp,
                                                 chunk_size=chunk, want_dense=want_dense,
                                                 right_padding=right_padding)
    loss = (alpha * ga.to(DEV)).sum()
    if se is not None:
        loss = loss + (beta * gb.to(DEV)).sum()
    loss.backward()
    return dense, alpha, beta, ppd.grad, (sed.grad if se is not None else None)


@pytest.mark.parametrize("promise", [False, True])
@pytest.mark.parametrize("name", list(FIXED))
def test_pooled_training_path_matches_reference_wrapper_goldens(name, promise):
    import simulst_b200
    c = FIXED[name]
    n, t, s, ratio, masked, soft, mp = [int(v) for v in c.cfg]
    se = c.soft_energy if soft else None
    dense, alpha, beta, gpp, gse = _run(c.p_pooled, s, ratio, se, opt(c.mask), bool(mp), c.g_alpha, c.g_beta,
                                        right_padding=promise)
    simulst_b200.check_status()
    assert torch.equal(dense.cpu(), c.p_choose)
    assert_parity(alpha, c.alpha, name + " alpha")
    assert_parity(beta, c.beta, name + " beta")
    floor = 4e-7 * s * float(max(c.g_alpha.abs().max(), c.g_beta.abs().max()))
    assert_parity(gpp, c.grad_p_pooled, name + " grad_p_pooled", extra_atol=floor)
    if soft:
        assert_parity(gse, c.grad_soft_energy, name + " grad_soft_energy", extra_atol=floor)


def _seeded(n, t, s, ratio, seed, masked):
    g = torch.Generator().manual_seed(seed)
    sp = (s + ratio - 1) // ratio
    pp = torch.sigmoid(torch.randn(n, t, sp, generator=g) - 1.0)
    se = torch.randn(n, t, s, generator=g)
    ga = torch.randn(n, t, s, generator=g) * 1e-2 + (torch.arange(s) + 1.0) / s
    gb = torch.randn(n, t, s, generator=g)
    mask = None
    if masked:
        lens = torch.randint(s // 2, s + 1, (n,), generator=g)
        lens[0] = s
        mask = torch.arange(s)[None, :] >= lens[:, None]
    return pp, se, ga, gb, mask


@pytest.mark.parametrize("n,t,s,ratio,masked,soft,dtype", [
    (2, 128, 1024, 8, False, True, torch.float32),       # training shape rows (BASELINE config 2), fused
    (2, 128, 1024, 8, False, True, torch.bfloat16),
    (2, 64, 1000, 8, True, True, torch.float32),         # tail fix-up + right padding (promise -> fused)
    (2, 64, 1000, 8, True, False, torch.bfloat16),       # hard-aligned
    (2, 32, 2048, 8, False, True, torch.float32),        # 256-thread CTAs
    (1, 16, 4096, 16, True, True, torch.float32),        # 512-thread CTAs
    (2, 16, 999, 8, False, True, torch.float32),         # rows not 16-byte multiples: expanded, dense path
    (2, 16, 1024, 4, False, True, torch.float32),        # ratio 4: two grid columns per thread
    (2, 16, 1024, 3, True, True, torch.float32),         # ratio 3: grid columns at uneven thread offsets
    (3, 9, 520, 16, True, False, torch.float32),         # hard-aligned, ratio 16, residual off the grid
    (1, 8, 6000, 8, False, True, torch.float32),         # rows beyond 4096 frames: expanded
    (2, 24, 512, 8, False, True, torch.float32),         # chunkwise: expanded (run with chunk below)
])
def test_pooled_training_path_matches_oracle(n, t, s, ratio, masked, soft, dtype):
    import simulst_b200
    pp, se, ga, gb, mask = _seeded(n, t, s, ratio, 31 + s + ratio, masked)
    pp, se = pp.to(dtype).float(), se.to(dtype).float()          # the oracle sees the rounded inputs
    chunk = 4 if (s == 512 and soft) else None
    dense, alpha, beta, gpp, gse = _run(pp, s, ratio, se if soft else None, mask, True, ga, gb, dtype=dtype,
                                        right_padding=masked, chunk=chunk)
    simulst_b200.check_status()
    ppo = pp.clone().requires_grad_()
    seo = se.clone().requires_grad_() if soft else None
    p_o, a_o, b_o = omma.mma_process_train_pooled(ppo, s, ratio, seo, mask, 1e-6, True, chunk)
    loss = (a_o * ga).sum()
    if soft:
        loss = loss + (b_o * gb).sum()
    loss.backward()
    p64 = pp.double().requires_grad_()
    s64 = se.double().requires_grad_() if soft else None
    _, a64, b64 = omma.mma_process_train_pooled(p64, s, ratio, s64, mask, 1e-6, True, chunk,
                                                compute_dtype=torch.float64)
    l64 = (a64 * ga.double()).sum()
    if soft:
        l64 = l64 + (b64 * gb.double()).sum()
    l64.backward()
    tag = f"pooled n{n} t{t} s{s} r{ratio} {dtype}"
    assert torch.equal(dense.float().cpu(), p_o.detach())
    assert_parity(alpha, a_o, tag + " alpha", a64)
    assert_parity(beta, b_o, tag + " beta", b64)
    floor = 4e-7 * s * float(max(ga.abs().max(), gb.abs().max()))
    rt = 1e-5 if dtype == torch.float32 else 2.0 ** -8
    assert_parity(gpp, ppo.grad, tag + " grad_p_pooled", p64.grad, rtol=rt, extra_atol=floor)
    if soft:
        assert_parity(gse, seo.grad, tag + " grad_energy", s64.grad, rtol=rt, extra_atol=floor)


@pytest.mark.parametrize("s,masked", [(1024, False), (1000, True), (264, True)])
def test_pooled_grid_kernels_agree_with_dense_kernels_on_the_expanded_row(s, masked):
    """The pooled-grid path (mma_sparse.cu) against the dense kernels fed the zero-upsampled tensor
    -- and the expand + dense route of the pooled entry points (simulst_mma_set_pooled_grid(0)),
    which must agree with the dense entry points bit for bit."""
    import simulst_b200
    from simulst_b200 import _lib, ops
    lib = _lib.load()
    n, t, ratio = 3, 20, 8
    pp, se, ga, gb, mask = _seeded(n, t, s, ratio, 77, masked)
    dense, alpha, beta, gpp, gse = _run(pp, s, ratio, se, mask, True, ga, gb, right_padding=masked)
    lib.simulst_mma_set_pooled_grid(0)
    try:
        dense_x, alpha_x, beta_x, gpp_x, gse_x = _run(pp, s, ratio, se, mask, True, ga, gb, right_padding=masked)
    finally:
        lib.simulst_mma_set_pooled_grid(1)
    simulst_b200.assume_right_padding(masked)
    try:
        pd = dense.detach().clone().requires_grad_()
        sed = se.to(DEV).requires_grad_()
        a2, b2 = ops.mma_train(pd, sed, mask.to(DEV) if masked else None, eps=1e-6, mass_preservation=True)
        ((a2 * ga.to(DEV)).sum() + (b2 * gb.to(DEV)).sum()).backward()
    finally:
        simulst_b200.assume_right_padding(False)
    cols = torch.arange(1, pp.shape[-1] + 1) * ratio - 1
    cols[-1] = s - 1
    assert torch.equal(dense, dense_x)
    assert torch.equal(alpha_x, a2) and torch.equal(beta_x, b2) and torch.equal(gse_x, sed.grad)
    assert torch.equal(gpp_x, pd.grad[:, :, cols.to(DEV)])
    floor = 4e-7 * s * float(max(ga.abs().max(), gb.abs().max()))
    assert_parity(alpha, a2, "grid vs dense alpha")
    assert_parity(beta, b2, "grid vs dense beta")
    assert_parity(gse, sed.grad, "grid vs dense grad_energy", extra_atol=floor)
    assert_parity(gpp, gpp_x, "grid vs dense grad_p_pooled", extra_atol=floor)


@pytest.mark.parametrize("s,masked", [(1024, False), (520, True)])
def test_pooled_latency_loss_mode_without_dense_alpha(s, masked):
    """want_alpha=False: the dense alpha is never written; alpha reaches the loss through beta and the
    expected delays only (mma_criterion.py:146-157).  Outputs and gradients against the oracle, which
    forms the delays from its dense alpha."""
    import simulst_b200
    from simulst_b200 import ops
    n, t, ratio = 3, 24, 8
    pp, se, _, gb, mask = _seeded(n, t, s, ratio, 91, masked)
    g = torch.Generator().manual_seed(92)
    gd = torch.randn(n, t, generator=g) / s
    ppd, sed = pp.to(DEV).requires_grad_(), se.to(DEV).requires_grad_()
    dense, alpha, beta, delays = ops.mma_train_pooled(ppd, s, ratio, sed, mask.to(DEV) if masked else None,
                                                      with_delays=True, want_dense=False, right_padding=masked,
                                                      want_alpha=False)
    assert dense is None and alpha is None
    ((beta * gb.to(DEV)).sum() + (delays * gd.to(DEV)).sum()).backward()
    simulst_b200.check_status()
    ppo, seo = pp.clone().requires_grad_(), se.clone().requires_grad_()
    _, a_o, b_o = omma.mma_process_train_pooled(ppo, s, ratio, seo, mask, 1e-6, True, None)
    d_o = omma.expected_delays(a_o)
    ((b_o * gb).sum() + (d_o * gd).sum()).backward()
    assert_parity(beta, b_o, "lean beta")
    assert_parity(delays, d_o, "lean delays", extra_atol=2e-6 * s)
    floor = 4e-7 * s * float(max(gb.abs().max(), (gd.abs().max() * s)))
    assert_parity(ppd.grad, ppo.grad, "lean grad_p_pooled", extra_atol=floor)
    assert_parity(sed.grad, seo.grad, "lean grad_energy", extra_atol=floor)


@pytest.mark.parametrize("s,masked", [(1024, False), (1000, True), (2048, False)])
def test_pooled_grid_kernels_are_deterministic(s, masked):
    """The recurrence kernels of the pooled-grid path are warp-specialised: producer, chain and consumer
    warps exchange a step's operands through shared-memory rings guarded by mbarriers.  A missing
    ordering edge would show up as run-to-run differences; 40 repetitions must be bit-identical."""
    n, t, ratio = 8, 96, 8
    pp, se, ga, gb, mask = _seeded(n, t, s, ratio, 123, masked)
    first = None
    for _ in range(40):
        out = _run(pp, s, ratio, se, mask, True, ga, gb, right_padding=masked)
        torch.cuda.synchronize()
        if first is None:
            first = [x.clone() for x in out]
        else:
            for a, b in zip(first, out):
                assert torch.equal(a, b)


def _random_pooled_cases(count, seed):
    g = torch.Generator().manual_seed(seed)
    cases = []
    for _ in range(count):
        ratio = [2, 3, 4, 8, 8, 8, 16][int(torch.randint(0, 7, (1,), generator=g))]
        s = int(torch.randint(1, 321, (1,), generator=g))
        if int(torch.randint(0, 2, (1,), generator=g)):
            s = max(4, s // 4 * 4)                    # half of the cases on the pooled-grid kernels (fp32: S % 4 == 0)
        t = int(torch.randint(1, 7, (1,), generator=g))
        n = int(torch.randint(1, 4, (1,), generator=g))
        masked, soft, mp = [bool(int(torch.randint(0, 2, (1,), generator=g))) for _ in range(3)]
        cases.append((n, t, s, ratio, masked, soft, mp))
    return cases


@pytest.mark.parametrize("n,t,s,ratio,masked,soft,mp", _random_pooled_cases(48, 2024))
def test_pooled_training_path_random_shapes(n, t, s, ratio, masked, soft, mp):
    """Seeded random geometry: source lengths 1..320 (shorter than the ratio, odd, multiples of 4 and 8),
    ratios 2..16, one to six target steps, with / without right padding (live lengths down to 1), soft
    attention and mass preservation -- every combination against the oracle, whichever kernels serve it."""
    import simulst_b200
    g = torch.Generator().manual_seed(n * 1000003 + t * 10007 + s * 101 + ratio)
    sp = (s + ratio - 1) // ratio
    pp = torch.sigmoid(torch.randn(n, t, sp, generator=g))
    se = torch.randn(n, t, s, generator=g) if soft else None
    ga = torch.randn(n, t, s, generator=g)
    gb = torch.randn(n, t, s, generator=g)
    mask = None
    if masked:
        lens = torch.randint(1, s + 1, (n,), generator=g)
        mask = torch.arange(s)[None, :] >= lens[:, None]
    dense, alpha, beta, gpp, gse = _run(pp, s, ratio, se, mask, mp, ga, gb, right_padding=masked)
    simulst_b200.check_status()
    ppo = pp.clone().requires_grad_()
    seo = se.clone().requires_grad_() if soft else None
    p_o, a_o, b_o = omma.mma_process_train_pooled(ppo, s, ratio, seo, mask, 1e-6, mp, None)
    loss = (a_o * ga).sum()
    if soft:
        loss = loss + (b_o * gb).sum()
    loss.backward()
    p64 = pp.double().requires_grad_()
    s64 = se.double().requires_grad_() if soft else None
    _, a64, b64 = omma.mma_process_train_pooled(p64, s, ratio, s64, mask, 1e-6, mp, None, compute_dtype=torch.float64)
    l64 = (a64 * ga.double()).sum()
    if soft:
        l64 = l64 + (b64 * gb.double()).sum()
    l64.backward()
    tag = f"random pooled n{n} t{t} s{s} r{ratio} m{int(masked)} soft{int(soft)} mp{int(mp)}"
    assert torch.equal(dense.cpu(), p_o.detach())
    assert_parity(alpha, a_o, tag + " alpha", a64)
    assert_parity(beta, b_o, tag + " beta", b64)
    floor = 4e-7 * s * float(max(ga.abs().max(), gb.abs().max()))
    assert_parity(gpp, ppo.grad, tag + " grad_p_pooled", p64.grad, extra_atol=floor)
    if soft:
        assert_parity(gse, seo.grad, tag + " grad_energy", s64.grad, extra_atol=floor)


def test_pooled_without_dense_output_and_is_fused_query():
    from simulst_b200 import _lib, ops
    lib = _lib.load()
    assert lib.simulst_mma_pooled_is_fused(_lib.BF16, 1024, 8, 0, _lib.MMA_SOFT | _lib.MMA_MASS_PRESERVATION, 0) == 1
    assert lib.simulst_mma_pooled_is_fused(_lib.BF16, 1024, 8, 0, _lib.MMA_SOFT, 1) == 0      # mask, no promise
    assert lib.simulst_mma_pooled_is_fused(_lib.BF16, 1024, 8, 0, _lib.MMA_SOFT | _lib.MMA_RIGHT_PADDING, 1) == 1
    assert lib.simulst_mma_pooled_is_fused(_lib.F32, 999, 8, 0, _lib.MMA_SOFT, 0) == 0
    assert lib.simulst_mma_pooled_is_fused(_lib.F32, 1024, 8, 4, _lib.MMA_SOFT, 0) == 0       # chunkwise
    pp, se, ga, gb, _ = _seeded(2, 6, 256, 8, 5, False)
    d0, a0, b0, g0, e0 = _run(pp, 256, 8, se, None, True, ga, gb, want_dense=True)
    d1, a1, b1, g1, e1 = _run(pp, 256, 8, se, None, True, ga, gb, want_dense=False)
    assert d1 is None and torch.equal(a0, a1) and torch.equal(b0, b1) and torch.equal(g0, g1) and torch.equal(e0, e1)
    # a non-fused call without the dense buffer is an argument error at the C ABI
    a = torch.empty(2, 6, 999, device=DEV)
    p = torch.rand(2, 6, 125, device=DEV)
    rc = lib.simulst_mma_train_fwd_pooled(_lib.ptr(p), _lib.F32, 8, None, 0, None, None, _lib.ptr(a), None, None,
                                          None, None, 2, 6, 999, 1e-6, 0, 0, None,
                                          _lib.stream_ptr(torch.device(DEV)))
    assert rc == -1


def test_pooled_bad_probabilities_and_broken_promise_are_reported():
    import simulst_b200
    from simulst_b200 import ops
    pp = torch.rand(2, 4, 32, device=DEV)
    pp[1, 2, 5] = 1.5
    ops.mma_train_pooled(pp, 256, 8, None, None)
    with pytest.raises(AssertionError):
        simulst_b200.check_status()
    pp = torch.rand(2, 4, 32, device=DEV)
    mask = torch.zeros(2, 256, dtype=torch.bool, device=DEV)
    mask[1, 100:120] = True                 # a hole: not a right-padding mask
    _, alpha, _, _ = ops.mma_train_pooled(pp, 256, 8, None, mask, right_padding=True)
    assert bool(torch.isnan(alpha[1]).all()) and not bool(torch.isnan(alpha[0]).any())
    with pytest.raises(RuntimeError):
        simulst_b200.check_status()


@pytest.mark.reference
@pytest.mark.parametrize("kind", ["infinite_lookback", "hard_aligned"])
@pytest.mark.parametrize("how", ["mixin", "patch"])
@pytest.mark.parametrize("s,masked", [(72, False), (77, True)])
def test_reference_fixed_pre_decision_class_forward(kind, how, s, masked):
    """The real `*_fixed_pre_decision` class (modules/fixed_pre_decision.py:175-190) with the
    training body replaced, through the reference's own forward() on the GPU, against the
    untouched class on the CPU: outputs, the attention dict and every parameter gradient."""
    if not ref_loader.available():
        pytest.skip("reference files not present")
    from simulst_b200.modules.fixed_pre_decision import B200FixedStrideMixin, patch_fixed_pre_decision
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        heads, embed, t, bsz = 4, 64, 9, 3
        ref = ref_loader.make_fixed_pre_decision_attention(kind, 8, "average", embed, heads, seed=11)
        ref.noise_std = 0.0
        ref.train()
        mine = copy.deepcopy(ref)
        base = type(ref)
        if how == "mixin":
            mine.__class__ = type("B200" + base.__name__, (B200FixedStrideMixin, base), {})
        else:
            mine.__class__ = patch_fixed_pre_decision(type("Patched" + base.__name__, (base,), {}))
        mine = mine.to(DEV)
        g = torch.Generator().manual_seed(12)
        q = torch.randn(t, bsz, embed, generator=g)
        k = torch.randn(s, bsz, embed, generator=g)
        mask = None
        if masked:
            lens = torch.randint(s // 2, s + 1, (bsz,), generator=g)
            lens[0] = s
            mask = torch.arange(s)[None, :] >= lens[:, None]
        out_r, extra_r = ref(q, k, k, key_padding_mask=mask)
        w_out = torch.randn(out_r.shape, generator=g)
        w_alpha = torch.randn(extra_r["alpha"].shape, generator=g) * 0.1
        (out_r * w_out).sum().add((extra_r["alpha"] * w_alpha).sum()).backward()
        out_m, extra_m = mine(q.to(DEV), k.to(DEV), k.to(DEV), key_padding_mask=mask.to(DEV) if masked else None)
        (out_m * w_out.to(DEV)).sum().add((extra_m["alpha"] * w_alpha.to(DEV)).sum()).backward()
        import simulst_b200
        simulst_b200.check_status()
        tol = dict(rtol=2e-4, atol=2e-5)        # cuBLAS vs MKL projections around the path
        torch.testing.assert_close(extra_m["p_choose"].cpu(), extra_r["p_choose"], **tol)
        torch.testing.assert_close(extra_m["alpha"].cpu(), extra_r["alpha"], **tol)
        torch.testing.assert_close(extra_m["beta"].cpu(), extra_r["beta"], **tol)
        torch.testing.assert_close(out_m.cpu(), out_r, **tol)
        for (name, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
            if pr.grad is None:
                assert pm.grad is None or float(pm.grad.abs().max()) == 0.0, name
                continue
            scale = float(pr.grad.abs().max())
            torch.testing.assert_close(pm.grad.cpu(), pr.grad, rtol=1e-3, atol=1e-4 * max(scale, 1e-3),
                                       msg=lambda m, name=name: f"{name}: {m}")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
