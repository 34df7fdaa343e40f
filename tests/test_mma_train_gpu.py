"""GPU parity of the fused MMA training kernels (forward and backward) against
(a) the committed golden vectors produced by the unmodified reference and
(b) the CPU oracle on seeded inputs, through the C ABI (ctypes) path.

Tolerance (tests/parity.py): rtol 1e-5, atol 1e-6 x the tensor's scale against the reference's
fp32 result; elements that miss that gate must be at least as close to the fp64
restatement as the reference itself is, and every use of that second assertion is reported.
"""
import pytest
import torch

from oracle import mma as omma
from tests.golden_io import load, opt
from tests.parity import assert_parity  # noqa: F401  (re-exported for the other GPU test modules)

pytestmark = pytest.mark.gpu

TRAIN = load("mma_train.npz")
RTOL = 1e-5


DEFAULT_PIPELINE = 5


@pytest.fixture(params=[5, 1, 0], ids=["default", "pipe-fwd+generic-bwd", "generic"])
def kernel_family(request):
    """simulst_mma_set_pipeline mode: default (pipelined forward + dense fast-path backward where
    the row qualifies), pipelined forward + generic backward, both generic -- every family has to
    pass the same parity gate."""
    from simulst_b200 import _lib
    lib = _lib.load()
    assert lib.simulst_mma_set_pipeline(request.param) == 0
    yield request.param
    lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE)


def _run(p, se, mask, mp, chunk, soft, g_alpha, g_beta, dtype=torch.float32):
    from simulst_b200 import ops
    dev = torch.device("cuda")
    p_d = p.to(dev, dtype).requires_grad_()
    se_d = se.to(dev, dtype).requires_grad_() if soft else None
    m_d = mask.to(dev) if mask is not None else None
    alpha, beta = ops.mma_train(p_d, se_d, m_d, eps=1e-6, mass_preservation=mp,
                                chunk_size=chunk or None)
    loss = (alpha * g_alpha.to(dev)).sum()
    if soft:
        loss = loss + (beta * g_beta.to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    return (alpha.detach().cpu(), beta.detach().cpu(), p_d.grad.float().cpu(),
            se_d.grad.float().cpu() if soft else None)


@pytest.mark.parametrize("name", list(TRAIN))
def test_mma_train_matches_reference_golden(name, kernel_family):
    c = TRAIN[name]
    n, t, s, masked, chunk, soft, mp = [int(v) for v in c.cfg]
    alpha, beta, gp, ge = _run(c.p, c.soft_energy, opt(c.mask), bool(mp), chunk, bool(soft),
                               c.g_alpha, c.g_beta)
    a64, b64, gp64, ge64 = _fp64(c.p, c.soft_energy if soft else None, opt(c.mask), bool(mp), chunk,
                                 c.g_alpha, c.g_beta)
    assert_parity(alpha, c.alpha, f"golden {name} alpha", a64)
    if soft:
        assert_parity(beta, c.beta, f"golden {name} beta", b64)
    assert_parity(gp, c.grad_p, f"golden {name} grad_p", gp64)
    if soft:
        assert_parity(ge, c.grad_soft_energy, f"golden {name} grad_soft_energy", ge64)


def _fp64(p, se, mask, mp, chunk, ga, gb):
    """fp64 restatement (forward + autograd) of the same formulas: the yardstick of the second
    parity assertion."""
    p64 = p.double().requires_grad_()
    se64 = se.double().requires_grad_() if se is not None else None
    a64, b64 = omma.mma_process_train(p64, se64, mask, 1e-6, mp, chunk or None, compute_dtype=torch.float64)
    loss = (a64 * ga).sum()
    if se is not None:
        loss = loss + (b64 * gb).sum()
    loss.backward()
    return a64.detach(), b64.detach(), p64.grad, (se64.grad if se is not None else None)


def _seeded(n, t, s, seed, mu=-2.0, masked=False):
    g = torch.Generator().manual_seed(seed)
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) + mu)
    se = torch.randn(n, t, s, generator=g)
    mask = None
    if masked:
        lens = torch.randint(max(1, s // 2), s + 1, (n,), generator=g)
        lens[0] = s
        mask = torch.arange(s)[None, :] >= lens[:, None]
    ga = (torch.arange(1, s + 1).float() / s).expand(n, t, s) + 1e-2 * torch.randn(n, t, s, generator=g)
    gb = torch.randn(n, t, s, generator=g)
    return p, se, mask, ga, gb


CASES = [
    # n, t, s, masked, chunk, config override (threads, vpt)
    (8, 32, 256, False, 0, None),          # BASELINE config 1 rows
    (4, 16, 1024, False, 0, (256, 8)),
    (4, 16, 1024, True, 0, (128, 8)),
    (3, 8, 1000, True, 0, None),           # S not a multiple of the vector width
    (3, 8, 999, False, 0, None),           # unaligned rows -> cooperative (non-TMA) staging
    (2, 6, 2048, False, 0, None),
    (2, 4, 6000, True, 0, None),           # long-form
    (2, 5, 300, True, 7, None),            # chunkwise
    (2, 12, 96, False, 0, (32, 8)),
]


@pytest.mark.parametrize("n,t,s,masked,chunk,cfg", CASES)
def test_mma_train_matches_oracle(n, t, s, masked, chunk, cfg, kernel_family):
    from simulst_b200 import _lib
    lib = _lib.load()
    p, se, mask, ga, gb = _seeded(n, t, s, seed=100 + s + t, masked=masked)
    if cfg is not None:
        assert lib.simulst_mma_set_config(*cfg) == 0
    try:
        alpha, beta, gp, ge = _run(p, se, mask, True, chunk, True, ga, gb)
    finally:
        lib.simulst_mma_set_config(0, 0)
    p_o = p.clone().requires_grad_()
    se_o = se.clone().requires_grad_()
    a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, True, chunk or None)
    ((a_o * ga).sum() + (b_o * gb).sum()).backward()
    p64 = p.double().requires_grad_()
    se64 = se.double().requires_grad_()
    a64, b64 = omma.mma_process_train(p64, se64, mask, 1e-6, True, chunk or None,
                                      compute_dtype=torch.float64)
    ((a64 * ga).sum() + (b64 * gb).sum()).backward()
    a64, b64 = a64.detach(), b64.detach()
    assert_parity(alpha, a_o.detach(), "alpha", a64)
    assert_parity(beta, b_o.detach(), "beta", b64)
    # gradients are length-S sums of upstream-gradient-sized terms: rounding floor 2*2^-24*sqrt(S)*|g|
    floor = 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(ga.abs().max()), float(gb.abs().max()))
    assert_parity(gp, p_o.grad, "grad_p", p64.grad, extra_atol=floor)
    assert_parity(ge, se_o.grad, "grad_soft_energy", se64.grad, extra_atol=floor)
    # accuracy against the fp64 restatement: not worse than 2x the reference's own error
    err_k = (alpha.double() - a64).abs().max().item()
    err_r = (a_o.detach().double() - a64).abs().max().item()
    assert err_k <= 2 * err_r + 1e-6, (err_k, err_r)


def test_tma_and_cooperative_staging_agree():
    """Generic kernels: rows staged by TMA bulk copies or by cooperative loads give the same bits."""
    from simulst_b200 import _lib
    lib = _lib.load()
    p, se, mask, ga, gb = _seeded(3, 9, 512, seed=5, masked=True)
    outs = []
    lib.simulst_mma_set_pipeline(0)
    try:
        for tma in (1, 0):
            lib.simulst_mma_set_tma(tma)
            outs.append(_run(p, se, mask, True, 0, True, ga, gb))
    finally:
        lib.simulst_mma_set_tma(1)
        lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE)
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape", [(3, 17, 1024), (3, 7, 1000), (2, 5, 1504), (2, 3, 6000), (2, 5, 264), (2, 4, 4096)])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("soft", [False, True])
def test_pipelined_and_generic_kernels_agree(masked, soft, shape):
    """The software-pipelined kernels and the generic (one scan per barrier) kernels evaluate the
    same formulas with different scan groupings: results agree to a few ulp of the row scale."""
    from simulst_b200 import _lib
    lib = _lib.load()
    p, se, mask, ga, gb = _seeded(*shape, seed=11, masked=masked)
    outs = []
    try:
        for pipe in (1, 0):
            lib.simulst_mma_set_pipeline(pipe)
            outs.append(_run(p, se if soft else None, mask, True, 0, soft, ga, gb if soft else None))
    finally:
        lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE)
    for a, b in zip(*outs):
        if a is None:
            assert b is None
            continue
        # rounding differences live at the scale of the incoming gradients (suffix sums of
        # gradient-sized terms that largely cancel), not at the scale of the result
        scale = max(float(b.abs().max()), float(ga.abs().max()), float(gb.abs().max()) if soft else 0.0)
        torch.testing.assert_close(a, b, rtol=5e-6, atol=5e-6 * scale)


@pytest.mark.parametrize("shape", [(3, 17, 1024), (5, 9, 256), (4, 6, 512), (2, 5, 2048), (2, 4, 4096), (2, 4, 6144),
                                   (3, 6, 1000), (2, 5, 1504), (2, 4, 6000), (2, 5, 264),
                                   (2, 1, 1024), (3, 2, 512), (2, 1, 6000), (150, 3, 256)])
@pytest.mark.parametrize("soft", [False, True])
@pytest.mark.parametrize("mp", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fast_backward_matches_generic(shape, soft, mp, dtype):
    """The dense fast-path backward (mma_bwd_fast.cuh) performs the generic kernel's arithmetic
    in the same order: bit-identical without mass preservation; with it, the correction
    ok*g'_last is formed from block totals (one rounding apart), so results agree to a few ulp.
    Shapes: rows that fill the CTA (4..16 warps, 8 and 12 elements per thread) and ragged rows
    (S a multiple of the per-thread element count, tails neutralised in shared memory)."""
    from simulst_b200 import _lib
    lib = _lib.load()
    n, t, s_len = shape
    p, se, _, ga, gb = _seeded(n, t, s_len, seed=21, masked=False)
    p, se = p.to(dtype), se.to(dtype)
    outs = []
    try:
        for pipe in (4, 0):
            lib.simulst_mma_set_pipeline(pipe)
            outs.append(_run(p, se if soft else None, None, mp, 0, soft, ga, gb if soft else None, dtype=dtype))
    finally:
        lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE)
    for k, (a, b) in enumerate(zip(*outs)):
        if a is None:
            assert b is None
            continue
        if not mp or k < 2:
            assert torch.equal(a, b), f"output {k}"
        else:
            # the generic kernel forms suffix(mz*P*g0) - ok*g'_last*suffix(mz*P) (two sums of the
            # size of the incoming gradient, then a cancellation); the fast kernel subtracts
            # first.  The difference is rounding at the scale of the incoming gradients.
            scale = max(float(b.float().abs().max()), float(ga.abs().max()), float(gb.abs().max()) if soft else 0.0)
            ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 5e-6     # one 16-bit rounding step apart at most
            torch.testing.assert_close(a.float(), b.float(), rtol=ulp, atol=2e-6 * scale)


@pytest.mark.parametrize("shape", [(4, 12, 512), (2, 3, 4096), (2, 3, 6000)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_inputs_match_oracle_fed_upcast_values(dtype, shape, kernel_family):
    """16-bit inputs at every row-capacity class (8, 8 and 12 elements per thread: the last one
    only allows 8-byte vector accesses for 2-byte elements)."""
    p, se, mask, ga, gb = _seeded(*shape, seed=9)
    p_h, se_h = p.to(dtype), se.to(dtype)
    alpha, beta, gp, ge = _run(p_h, se_h, None, True, 0, True, ga, gb, dtype=dtype)
    p_o = p_h.float().requires_grad_()
    se_o = se_h.float().requires_grad_()
    a_o, b_o = omma.mma_process_train(p_o, se_o, None, 1e-6, True, None)
    ((a_o * ga).sum() + (b_o * gb).sum()).backward()
    assert_parity(alpha, a_o.detach(), "alpha")
    assert_parity(beta, b_o.detach(), "beta")
    # gradients are rounded to the 16-bit input dtype on store
    half_eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    torch.testing.assert_close(gp, p_o.grad, rtol=half_eps, atol=1e-5 * float(p_o.grad.abs().max()))
    torch.testing.assert_close(ge, se_o.grad, rtol=half_eps, atol=1e-5 * float(se_o.grad.abs().max()))


def test_status_word_reports_bad_probabilities():
    import simulst_b200
    from simulst_b200 import ops
    dev = torch.device("cuda")
    p = torch.rand(2, 3, 64, device=dev)
    p[1, 2, 5] = 1.5
    ops.mma_train(p, None, None, mass_preservation=False)
    with pytest.raises(AssertionError, match="Incorrect values"):
        simulst_b200.check_status(dev)
    p[1, 2, 5] = float("nan")
    ops.mma_train(p, None, None, mass_preservation=False)
    with pytest.raises(AssertionError, match="Nan"):
        simulst_b200.check_status(dev)
    simulst_b200.check_status(dev)      # cleared


@pytest.mark.parametrize("chunks,streams", [(1, 1), (3, 2), (8, 2)])
def test_host_pipeline_matches_device_path(chunks, streams):
    """MMAHostPipeline (pinned host buffers, row chunks over three streams) returns bit-identical
    alpha / beta / gradients to one device-resident fused call: rows never interact."""
    from simulst_b200 import ops
    from simulst_b200.host_pipeline import MMAHostPipeline
    dev = torch.device("cuda")
    n, t, s = 20, 16, 256
    g = torch.Generator().manual_seed(77)
    dt = torch.bfloat16
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2.0).to(dt)
    e = torch.randn(n, t, s, generator=g).to(dt)
    ga = torch.randn(n, t, s, generator=g).to(dev)
    gb = torch.randn(n, t, s, generator=g).to(dev)
    p_d = p.to(dev).requires_grad_()
    e_d = e.to(dev).requires_grad_()
    alpha, beta = ops.mma_train(p_d, e_d, None, eps=1e-6, mass_preservation=True)
    ((alpha * ga).sum() + (beta * gb).sum()).backward()
    pipe = MMAHostPipeline(n, t, s, dtype=dt, device=dev, chunks=chunks, compute_streams=streams)
    p_h, e_h = p.pin_memory(), e.pin_memory()
    gp_h = torch.empty_like(p).pin_memory()
    ge_h = torch.empty_like(e).pin_memory()
    for _ in range(2):      # second call exercises buffer reuse across steps
        a2, b2 = pipe.step(p_h, e_h, ga, gb, gp_h, ge_h)
    torch.cuda.synchronize()
    assert torch.equal(a2, alpha.detach()) and torch.equal(b2, beta.detach())
    assert torch.equal(gp_h, p_d.grad.cpu()) and torch.equal(ge_h, e_d.grad.cpu())
    with pytest.raises(ValueError):
        pipe.step(p, e_h, ga, gb, gp_h, ge_h)       # unpinned host tensor


@pytest.mark.parametrize("shape", [(4, 9, 1024), (3, 7, 1000), (2, 5, 1504), (2, 4, 4096), (3, 6, 264), (2, 3, 6000)])
@pytest.mark.parametrize("soft", [False, True])
@pytest.mark.parametrize("mp", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_right_padding_promise_matches_arbitrary_mask_kernel(shape, soft, mp, dtype):
    """With simulst_b200.assume_right_padding(True) masked rows run the dense backward kernel
    (per-row live length instead of per-element mask tests): same gradients as the arbitrary-mask
    kernel on right-padded batches, zeros at padded columns, and parity with the oracle."""
    import simulst_b200
    from simulst_b200 import _lib
    lib = _lib.load()
    n, t, s_len = shape
    p, se, mask, ga, gb = _seeded(n, t, s_len, seed=31, masked=True)
    p, se = p.to(dtype), se.to(dtype)
    outs = []
    try:
        for promise in (True, False):
            simulst_b200.assume_right_padding(promise)
            lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE if promise else 0)
            outs.append(_run(p, se if soft else None, mask, mp, 0, soft, ga, gb if soft else None, dtype=dtype))
    finally:
        simulst_b200.assume_right_padding(False)
        lib.simulst_mma_set_pipeline(DEFAULT_PIPELINE)
    scale = max(float(ga.abs().max()), float(gb.abs().max()) if soft else 0.0)
    ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 5e-6
    for k, (a, b) in enumerate(zip(*outs)):
        if a is None:
            assert b is None
            continue
        if k >= 2:      # gradients: exactly zero where the mask is set
            assert not bool(a[mask.unsqueeze(1).expand_as(a)].any()), f"output {k}: non-zero gradient at padded columns"
        # 16-bit outputs: two roundings of slightly different fp32 values can land two steps apart
        atol = (2e-5 if dtype == torch.bfloat16 else 2e-6) * max(scale, float(b.float().abs().max()))
        torch.testing.assert_close(a.float(), b.float(), rtol=ulp, atol=atol)
    if dtype == torch.float32:
        p_o = p.clone().requires_grad_()
        se_o = se.clone().requires_grad_() if soft else None
        a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, mp, None)
        loss = (a_o * ga).sum() + ((b_o * gb).sum() if soft else 0.0)
        loss.backward()
        assert_parity(outs[0][2], p_o.grad, "grad_p", extra_atol=2e-6 * scale)
        if soft:
            assert_parity(outs[0][3], se_o.grad, "grad_soft_energy", extra_atol=2e-6 * scale)


def test_broken_right_padding_promise_is_reported():
    """The promise is verified by the forward kernel: a mask with a hole raises at the next status check."""
    import simulst_b200
    from simulst_b200 import ops
    dev = torch.device("cuda")
    p, se, mask, _, _ = _seeded(3, 4, 512, seed=41, masked=True)
    mask[1, 5] = True          # a padded column in the middle of row 1
    mask[1, -1] = False
    simulst_b200.check_status(dev)
    try:
        simulst_b200.assume_right_padding(True)
        alpha, beta = ops.mma_train(p.to(dev), se.to(dev), mask.to(dev))
        # the offending row is poisoned (cannot train on silently), the others are computed
        assert bool(torch.isnan(alpha[1]).all()) and bool(torch.isnan(beta[1]).all())
        assert not bool(torch.isnan(alpha[0]).any()) and not bool(torch.isnan(alpha[2]).any())
        with pytest.raises(RuntimeError, match="right-padding"):
            simulst_b200.check_status(dev)
        simulst_b200.check_status(dev)      # cleared
    finally:
        simulst_b200.assume_right_padding(False)


@pytest.mark.parametrize("soft", [False, True])
@pytest.mark.parametrize("mp", [False, True])
def test_masked_call_is_split_by_row_between_dense_and_general_kernels(soft, mp):
    """Default handling of a padding mask: rows whose mask is a right-padding mask run through the
    dense kernels, rows with any other mask through the arbitrary-mask kernels (each CTA
    classifies its own row).  A batch mixing both kinds must give what a single pass through the
    arbitrary-mask kernels gives, and match the oracle."""
    from simulst_b200 import _lib
    lib = _lib.load()
    n, t, s_len = 8, 6, 1024
    p, se, mask, ga, gb = _seeded(n, t, s_len, seed=51, masked=True)
    g = torch.Generator().manual_seed(52)
    holes = torch.rand(n, s_len, generator=g) < 0.1
    holes[:4] = False                    # rows 0..3 stay right-padded, rows 4..7 get holes
    holes[:, 0] = False
    mask = mask | holes
    outs = []
    try:
        for split in (1, 0):
            assert lib.simulst_mma_set_mask_split(split) == 0
            outs.append(_run(p, se if soft else None, mask, mp, 0, soft, ga, gb if soft else None))
    finally:
        lib.simulst_mma_set_mask_split(1)
    scale = max(float(ga.abs().max()), float(gb.abs().max()) if soft else 0.0)
    for k, (a, b) in enumerate(zip(*outs)):
        if a is None:
            assert b is None
            continue
        assert torch.equal(a[4:], b[4:]), f"output {k}: rows with holes must come from the same kernel"
        torch.testing.assert_close(a[:4], b[:4], rtol=5e-6, atol=2e-6 * max(scale, float(b.abs().max())))
    if not mp:      # with mass preservation the reference indexes src_len - 1, meaningless for masks with holes
        p_o = p.clone().requires_grad_()
        se_o = se.clone().requires_grad_() if soft else None
        a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, mp, None)
        ((a_o * ga).sum() + ((b_o * gb).sum() if soft else 0.0)).backward()
        assert_parity(outs[0][0], a_o.detach(), "alpha")
        assert_parity(outs[0][2], p_o.grad, "grad_p", extra_atol=2e-6 * scale)
