"""GPU parity of the MMA training path on source lengths whose rows are NOT 16-byte multiples
(S = 1500 in bf16, S = 999, odd S in 16-bit types) and / or do not divide among the threads of a CTA.

The Python wrapper allocates 16-byte pitched outputs for such shapes (simulst_b200.set_pitched_outputs,
default on) and the C ABI's `_pitched` entry points route them to the SHIFT instantiations of the dense
kernels (aligned-superset bulk copies, reads at the row's byte offset, a live length per row).  Every
case here is compared with the CPU oracle (= the reference's primitive sequence) under the shared
parity gate, in every kernel family, with and without a right-padding mask, and once more with dense
outputs (generic kernels) -- the two allocations must both pass.
"""
import pytest
import torch

import simulst_b200
from oracle import mma as omma
from tests.parity import assert_parity
from tests.test_mma_train_gpu import DEFAULT_PIPELINE, _seeded, kernel_family  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle(p, se, mask, mp, ga, gb, dt=torch.float32):
    p_o = p.detach().to(dt).clone().requires_grad_()
    se_o = se.detach().to(dt).clone().requires_grad_() if se is not None else None
    a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, mp, None, compute_dtype=dt)
    loss = (a_o * ga).sum()
    if se is not None:
        loss = loss + (b_o * gb).sum()
    loss.backward()
    return a_o.detach(), b_o.detach(), p_o.grad, (se_o.grad if se is not None else None)


def _run(p, se, mask, mp, ga, gb, dtype, delays=False):
    from simulst_b200 import ops
    p_d = p.to(DEV, dtype).requires_grad_()
    se_d = se.to(DEV, dtype).requires_grad_() if se is not None else None
    m_d = mask.to(DEV) if mask is not None else None
    if delays:
        alpha, beta, d = ops.mma_train_with_delays(p_d, se_d, m_d, eps=1e-6, mass_preservation=mp)
    else:
        alpha, beta = ops.mma_train(p_d, se_d, m_d, eps=1e-6, mass_preservation=mp)
        d = None
    loss = (alpha * ga.to(DEV)).sum()
    if se is not None:
        loss = loss + (beta * gb.to(DEV)).sum()
    loss.backward()
    torch.cuda.synchronize()
    return (alpha.detach(), beta.detach(), p_d.grad.float().cpu(),
            se_d.grad.float().cpu() if se is not None else None, d)


def _check(n, t, s, dtype, masked, soft, mp, tag):
    p, se, mask, ga, gb = _seeded(n, t, s, seed=4200 + 3 * s + t, masked=masked)
    p, se = p.to(dtype), se.to(dtype)
    se_in = se if soft else None
    a_o, b_o, gp_o, ge_o = _oracle(p.float(), se_in.float() if soft else None, mask, mp, ga, gb)
    a64, b64, gp64, ge64 = _oracle(p.float(), se_in.float() if soft else None, mask, mp, ga, gb, torch.float64)
    alpha, beta, gp, ge, _ = _run(p, se_in, mask, mp, ga, gb, dtype)
    assert tuple(alpha.shape) == (n, t, s)
    assert_parity(alpha, a_o, tag + " alpha", a64)
    if soft:
        assert_parity(beta, b_o, tag + " beta", b64)
    floor = 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(ga.abs().max()), float(gb.abs().max()))
    rt = 1e-5 if dtype == torch.float32 else 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert_parity(gp, gp_o, tag + " grad_p", gp64, rtol=rt, extra_atol=floor)
    if soft:
        assert_parity(ge, ge_o, tag + " grad_energy", ge64, rtol=rt, extra_atol=floor)
    return alpha


UNALIGNED = [
    # n, T, S, dtype
    (3, 12, 999, torch.float32),         # rows shifted by 0 / 4 / 8 / 12 bytes
    (3, 12, 1001, torch.bfloat16),       # odd S in a 16-bit type: rows start on odd elements (byte permute)
    (2, 128, 1500, torch.bfloat16),      # the CIF config's own S at the headline depth
    (2, 16, 1500, torch.float32),        # aligned rows, S % 8 = 4: the row ends inside a thread
    (4, 9, 37, torch.float16),           # one warp, 4 elements per thread
    (4, 9, 130, torch.bfloat16),
    (3, 10, 250, torch.float32),
    (2, 8, 1017, torch.bfloat16),        # needs the next CTA size for its 16 spare columns
    (2, 8, 2047, torch.float16),
    (2, 6, 3001, torch.bfloat16),
    (1, 6, 5003, torch.float32),         # 12 elements per thread
]


@pytest.mark.parametrize("n,t,s,dtype", UNALIGNED, ids=lambda v: str(v).replace("torch.", ""))
@pytest.mark.parametrize("masked", [False, True], ids=["nomask", "rightpad"])
def test_unaligned_rows_match_oracle(n, t, s, dtype, masked, kernel_family):
    _check(n, t, s, dtype, masked, True, True, f"unaligned n{n} T{t} S{s} {str(dtype)[6:]} m{int(masked)} pipe{kernel_family}")


@pytest.mark.parametrize("soft,mp", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("s,dtype", [(1500, torch.bfloat16), (999, torch.float32), (1001, torch.float16)])
def test_unaligned_rows_other_modes(s, dtype, soft, mp):
    _check(3, 10, s, dtype, False, soft, mp, f"unaligned-modes S{s} {str(dtype)[6:]} soft{int(soft)} mp{int(mp)}")
    _check(3, 10, s, dtype, True, soft, mp, f"unaligned-modes masked S{s} {str(dtype)[6:]} soft{int(soft)} mp{int(mp)}")


def test_right_padding_promise_on_unaligned_rows():
    """With the promise the masked call is ONE pass of the SHIFT kernel; results as without it."""
    simulst_b200.assume_right_padding(True)
    try:
        _check(5, 12, 1500, torch.bfloat16, True, True, True, "unaligned promise S1500 bf16")
        _check(5, 12, 999, torch.float32, True, True, True, "unaligned promise S999 f32")
        simulst_b200.check_status()
    finally:
        simulst_b200.assume_right_padding(False)


def test_dense_outputs_still_served():
    """set_pitched_outputs(False): contiguous outputs, generic kernels, same parity gate."""
    simulst_b200.set_pitched_outputs(False)
    try:
        a = _check(2, 10, 1500, torch.bfloat16, False, True, True, "dense-out S1500 bf16")
        assert a.is_contiguous()
        a = _check(2, 10, 999, torch.float32, True, True, True, "dense-out S999 f32 masked")
        assert a.is_contiguous()
    finally:
        simulst_b200.set_pitched_outputs(True)
    a = _check(2, 10, 1500, torch.bfloat16, False, True, True, "pitched-out S1500 bf16")
    assert a.stride(1) % 8 == 0 and a.stride(1) >= 1500 and a.stride(0) == 10 * a.stride(1)


def test_expected_delays_on_unaligned_rows():
    n, t, s = 3, 11, 1500
    for masked in (False, True):
        p, se, mask, ga, gb = _seeded(n, t, s, seed=99, masked=masked)
        alpha, beta, _, _, d = _run(p.bfloat16(), se.bfloat16(), mask, True, ga, gb, torch.bfloat16, delays=True)
        steps = torch.arange(1, s + 1, device=DEV, dtype=torch.float32)
        want = (alpha * steps).sum(-1)
        torch.testing.assert_close(d, want, rtol=2e-5, atol=2e-4)


def test_pitched_c_abi_inputs_and_padding_columns():
    """Direct C-ABI call with PITCHED INPUTS as well (x_padded[..., :S] views, odd base offsets): results
    equal the dense call's, the outputs' padding columns hold zeros and nothing is written past a row's
    pitch (sentinel check)."""
    from simulst_b200 import _lib
    lib = _lib.load()
    n, t, s = 3, 7, 1001
    dt = torch.bfloat16
    p, se, _, ga, gb = _seeded(n, t, s, seed=5)
    ld_in = 1013                                    # odd pitch: every row at another byte offset
    ld_out = int(lib.simulst_mma_out_pitch(s))
    assert ld_out % 8 == 0 and ld_out >= s
    pb = torch.full((n, t, ld_in), 0.5, dtype=dt, device=DEV)
    eb = torch.full((n, t, ld_in), 9.0, dtype=dt, device=DEV)
    pb[..., :s] = p.to(DEV, dt)
    eb[..., :s] = se.to(DEV, dt)
    sent = 1024.0
    alpha = torch.full((n, t, ld_out + 8), sent, device=DEV)
    beta = torch.full((n, t, ld_out + 8), sent, device=DEV)
    # pitch ld_out + 8 with 8 guard columns per row: [S, ld_out) may receive zeros, [ld_out, ld_out+8) nothing
    side = torch.zeros(n, t, 2, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    flags = _lib.MMA_SOFT | _lib.MMA_MASS_PRESERVATION
    st = _lib.stream_ptr(torch.device(DEV))
    rc = lib.simulst_mma_train_fwd_pitched(pb.data_ptr(), _lib.BF16, ld_in, eb.data_ptr(), _lib.BF16, ld_in, None,
                                           alpha.data_ptr(), ld_out + 8, beta.data_ptr(), ld_out + 8,
                                           side.data_ptr(), None, n, t, s, 1e-6, 0, flags, status.data_ptr(), st)
    assert rc == 0
    gab = torch.full((n, t, ld_in), 7.0, device=DEV)
    gbb = torch.full((n, t, ld_in), 7.0, device=DEV)
    gab[..., :s] = ga.to(DEV)
    gbb[..., :s] = gb.to(DEV)
    gp = torch.full((n, t, ld_out + 8), sent, dtype=dt, device=DEV)
    ge = torch.full((n, t, ld_out + 8), sent, dtype=dt, device=DEV)
    rc = lib.simulst_mma_train_bwd_pitched(pb.data_ptr(), _lib.BF16, ld_in, eb.data_ptr(), _lib.BF16, ld_in, None,
                                           alpha.data_ptr(), ld_out + 8, side.data_ptr(),
                                           gab.data_ptr(), ld_in, gbb.data_ptr(), ld_in, None,
                                           gp.data_ptr(), _lib.BF16, ld_out + 8, ge.data_ptr(), _lib.BF16, ld_out + 8,
                                           n, t, s, 1e-6, 0, flags, st)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    a_o, b_o, gp_o, ge_o = _oracle(p.to(dt).float(), se.to(dt).float(), None, True, ga, gb)
    a64, b64, gp64, ge64 = _oracle(p.to(dt).float(), se.to(dt).float(), None, True, ga, gb, torch.float64)
    assert_parity(alpha[..., :s], a_o, "pitched abi alpha", a64)
    assert_parity(beta[..., :s], b_o, "pitched abi beta", b64)
    floor = 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(ga.abs().max()), float(gb.abs().max()))
    assert_parity(gp[..., :s].float(), gp_o, "pitched abi grad_p", gp64, rtol=2.0 ** -8, extra_atol=floor)
    assert_parity(ge[..., :s].float(), ge_o, "pitched abi grad_energy", ge64, rtol=2.0 ** -8, extra_atol=floor)
    for name, x in (("alpha", alpha), ("beta", beta), ("grad_p", gp.float()), ("grad_energy", ge.float())):
        pad = x[..., s:ld_out]
        assert bool(((pad == 0) | (pad == sent)).all()), f"{name}: padding columns hold neither zeros nor the sentinel"
        assert bool((x[..., ld_out:] == sent).all()), f"{name}: wrote past the row pitch"
    # a pitch below S is a shape error
    assert lib.simulst_mma_train_fwd_pitched(pb.data_ptr(), _lib.BF16, s - 1, eb.data_ptr(), _lib.BF16, ld_in, None,
                                             alpha.data_ptr(), ld_out, beta.data_ptr(), ld_out, side.data_ptr(), None,
                                             n, t, s, 1e-6, 0, flags, status.data_ptr(), st) == -2


def test_prob_check_on_unaligned_rows():
    """prob_check (functions.py:9-17) through the status word on the SHIFT path: clean inputs leave it
    zero (the staged over-read beyond a row must not trip it), a probability above 1 raises."""
    from simulst_b200 import ops
    n, t, s = 2, 6, 1500
    p, se, _, _, _ = _seeded(n, t, s, seed=3)
    simulst_b200.check_status()
    # rows followed in memory by garbage that would fail the check if the over-read leaked into it
    buf = torch.full((n * t * s + 64,), 7.0, dtype=torch.bfloat16, device=DEV)
    pv = buf[: n * t * s].view(n, t, s)
    pv.copy_(p.to(DEV, torch.bfloat16))
    ops.mma_train(pv, se.to(DEV, torch.bfloat16), None)
    simulst_b200.check_status()
    bad = p.clone()
    bad[1, 3, 1499] = 1.5
    ops.mma_train(bad.to(DEV, torch.bfloat16), se.to(DEV, torch.bfloat16), None)
    with pytest.raises(AssertionError):
        simulst_b200.check_status()


def test_broken_promise_on_unaligned_rows_is_flagged():
    from simulst_b200 import ops
    n, t, s = 2, 5, 999
    p, se, _, _, _ = _seeded(n, t, s, seed=8)
    mask = torch.zeros(n, s, dtype=torch.bool)
    mask[1, 100:200] = True                         # a hole: not a right-padding mask
    simulst_b200.check_status()
    simulst_b200.assume_right_padding(True)
    try:
        alpha, _ = ops.mma_train(p.to(DEV), se.to(DEV), mask.to(DEV))
        assert bool(torch.isnan(alpha[1]).all())
        with pytest.raises(RuntimeError):
            simulst_b200.check_status()
    finally:
        simulst_b200.assume_right_padding(False)


def _random_cases(count, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    dts = [torch.float32, torch.bfloat16, torch.float16]
    for q in range(count):
        s = int(torch.randint(1, 4200, (1,), generator=g))
        if q % 5 == 0:          # capacities' edges: a row that leaves exactly 16 / 15 spare columns, tiny rows
            s = [1008, 1009, 2032, 2033, 3, 17, 4080, 4081][(q // 5) % 8]
        n = int(torch.randint(2, 6, (1,), generator=g))
        t = int(torch.randint(1, 9, (1,), generator=g))
        out.append((n, t, s, dts[q % 3], q % 4, bool((q // 2) % 2), bool((q // 3) % 2) or q % 4 == 0))
    return out


@pytest.mark.parametrize("n,t,s,dtype,mask_kind,soft,mp", _random_cases(48, 2024),
                         ids=lambda v: str(v).replace("torch.", ""))
def test_random_lengths_and_masks_match_oracle(n, t, s, dtype, mask_kind, soft, mp):
    """Seeded random source lengths (unaligned rows and rows at the edges of a CTA's capacity), dtypes and
    modes; mask_kind 0 none, 1 right-padded with random lengths down to 1 (split by row), 2 the same with the
    right-padding promise, 3 right-padded lengths plus one arbitrary (holey) mask row."""
    g = torch.Generator().manual_seed(31 * s + n)
    p, se, _, ga, gb = _seeded(n, t, s, seed=12000 + s)
    p, se = p.to(dtype), se.to(dtype)
    mask = None
    if mask_kind:
        lens = torch.randint(1, s + 1, (n,), generator=g)
        lens[0] = s
        mask = torch.arange(s)[None, :] >= lens[:, None]
        if mask_kind == 3 and s > 4:
            mask[-1] = False
            mask[-1, 1] = True          # a hole: this row takes the arbitrary-mask pass
    se_in = se if soft else None
    a_o, b_o, gp_o, ge_o = _oracle(p.float(), se.float() if soft else None, mask, mp, ga, gb)
    a64, b64, gp64, ge64 = _oracle(p.float(), se.float() if soft else None, mask, mp, ga, gb, torch.float64)
    simulst_b200.assume_right_padding(mask_kind == 2)
    try:
        alpha, beta, gp, ge, _ = _run(p, se_in, mask, mp, ga, gb, dtype)
        simulst_b200.check_status()
    finally:
        simulst_b200.assume_right_padding(False)
    tag = f"random n{n} T{t} S{s} {str(dtype)[6:]} mask{mask_kind} soft{int(soft)} mp{int(mp)}"
    assert_parity(alpha, a_o, tag + " alpha", a64)
    if soft:
        assert_parity(beta, b_o, tag + " beta", b64)
    floor = 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(ga.abs().max()), float(gb.abs().max()))
    rt = 1e-5 if dtype == torch.float32 else 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert_parity(gp, gp_o, tag + " grad_p", gp64, rtol=rt, extra_atol=floor)
    if soft:
        assert_parity(ge, ge_o, tag + " grad_energy", ge64, rtol=rt, extra_atol=floor)


def test_rows_without_live_columns():
    """A padding mask that leaves a row no live column: alpha, beta and both gradients of that row are zero
    (the reference would fail on such a row with a scatter index of -1); its neighbours are unaffected."""
    from simulst_b200 import ops
    n, t, s = 3, 6, 1500
    p, se, _, ga, gb = _seeded(n, t, s, seed=77)
    lens = torch.tensor([s, 0, 733])
    mask = torch.arange(s)[None, :] >= lens[:, None]
    for promise in (False, True):
        simulst_b200.assume_right_padding(promise)
        try:
            alpha, beta, gp, ge, _ = _run(p.bfloat16(), se.bfloat16(), mask, True, ga, gb, torch.bfloat16)
        finally:
            simulst_b200.assume_right_padding(False)
        for x in (alpha[1], beta[1], gp[1], ge[1]):
            assert float(x.abs().max()) == 0.0
        keep = torch.tensor([0, 2])
        a_o, b_o, gp_o, ge_o = _oracle(p.bfloat16().float()[keep], se.bfloat16().float()[keep], mask[keep], True,
                                       ga[keep], gb[keep])
        assert_parity(alpha[keep], a_o, "empty-row neighbours alpha")
        assert_parity(beta[keep], b_o, "empty-row neighbours beta")
