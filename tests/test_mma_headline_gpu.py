"""GPU parity at the shapes the numbers are quoted on (BASELINE config 2 and the corners of the
config 5 sweep), CUDA-graph capture of the C-ABI calls, and mass_preservation(left_padding=True).

The T-step recurrence compounds rounding, so parity at T = 17 says little about T = 128 or 512;
these cases run the full depth against the CPU oracle (fp32 restatement = the reference's
primitive sequence, fp64 restatement as the yardstick of the second assertion, tests/parity.py).
"""
import pytest
import torch

from oracle import mma as omma
from tests.parity import assert_parity
from tests.test_mma_train_gpu import DEFAULT_PIPELINE, _run, _seeded, kernel_family  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
DEV = "cuda"

HEADLINE = [
    # n, T, S, dtype                       what
    (2, 128, 1024, torch.float32),       # BASELINE config 2 rows, fp32
    (2, 128, 1024, torch.bfloat16),      # BASELINE config 2 rows as benched (bf16 in, fp32 accumulate)
    (2, 512, 512, torch.float32),        # config 5: deepest recurrence
    (1, 256, 6000, torch.float32),       # config 5: longest rows
    (2, 128, 1500, torch.bfloat16),      # S*2 bytes not a multiple of 16 (the CIF config's own S)
    (2, 64, 1504, torch.float32),        # S between two CTA sizes
    # CTA sizes in one-warp steps (csrc/mma_dispatch.h): 96, 160, 192, 224, 320, 384, 448 threads
    (2, 24, 768, torch.bfloat16),
    (2, 24, 1280, torch.float32),
    (2, 24, 1536, torch.bfloat16),
    (2, 24, 1792, torch.float32),
    (2, 16, 2560, torch.bfloat16),
    (2, 16, 3000, torch.float32),        # ragged inside a 384-thread CTA
    (2, 16, 3584, torch.bfloat16),
    (3, 16, 600, torch.float32),         # ragged inside a 96-thread CTA
]


def _oracle(p, se, mask, ga, gb, dtype64=False):
    dt = torch.float64 if dtype64 else torch.float32
    p_o = p.detach().to(dt).clone().requires_grad_()
    se_o = se.detach().to(dt).clone().requires_grad_()
    a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, True, None, compute_dtype=dt)
    ((a_o * ga).sum() + (b_o * gb).sum()).backward()
    return a_o.detach(), b_o.detach(), p_o.grad, se_o.grad


def grad_floor(s, *upstream):
    """Absolute rounding floor of a gradient: every d/dp, d/dE is a length-S prefix / suffix sum of
    terms of the size of the upstream gradients, so ANY fp32 evaluation (the reference's included)
    carries ~2^-24 * sqrt(S) * |g| of accumulated rounding whatever the size of the result."""
    return 2.0 * 2.0 ** -24 * s ** 0.5 * max(float(g.abs().max()) for g in upstream)


_ORACLE_CACHE = {}


@pytest.mark.parametrize("n,t,s,dtype", HEADLINE, ids=lambda v: str(v).replace("torch.", ""))
def test_headline_shapes_match_oracle(n, t, s, dtype, kernel_family):
    key = (n, t, s, dtype)
    p, se, _, ga, gb = _seeded(n, t, s, seed=7000 + s + t)
    p, se = p.to(dtype), se.to(dtype)           # the oracle is fed the same (rounded) values, up-cast
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = (_oracle(p.float(), se.float(), None, ga, gb),
                              _oracle(p.float(), se.float(), None, ga, gb, dtype64=True))
    (a_o, b_o, gp_o, ge_o), (a64, b64, gp64, ge64) = _ORACLE_CACHE[key]
    alpha, beta, gp, ge = _run(p, se, None, True, 0, True, ga, gb, dtype=dtype)
    tag = f"headline n{n} T{t} S{s} {str(dtype)[6:]} pipe{kernel_family}"
    assert_parity(alpha, a_o, tag + " alpha", a64)
    assert_parity(beta, b_o, tag + " beta", b64)
    floor = grad_floor(s, ga, gb)
    if dtype == torch.float32:
        assert_parity(gp, gp_o, tag + " grad_p", gp64, extra_atol=floor)
        assert_parity(ge, ge_o, tag + " grad_energy", ge64, extra_atol=floor)
    else:
        # gradients are rounded to bf16 on store: one rounding step of the 16-bit type on top of the gate
        half = 2.0 ** -8
        assert_parity(gp, gp_o, tag + " grad_p", gp64, rtol=half, extra_atol=floor)
        assert_parity(ge, ge_o, tag + " grad_energy", ge64, rtol=half, extra_atol=floor)


# ----------------------------------------------------------------------------- CUDA graphs
def _raw_train_call(lib, _lib, bufs, n, t, s, flags):
    st = _lib.stream_ptr(torch.device(DEV))
    p, e, alpha, beta, side, ga, gb, gp, ge, status = bufs
    rc = lib.simulst_mma_train_fwd(_lib.ptr(p), _lib.dtype_enum(p.dtype), _lib.ptr(e), _lib.dtype_enum(e.dtype),
                                   None, _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(side), n, t, s, 1e-6, 0, flags,
                                   _lib.ptr(status), st)
    assert rc == 0
    rc = lib.simulst_mma_train_bwd(_lib.ptr(p), _lib.dtype_enum(p.dtype), _lib.ptr(e), _lib.dtype_enum(e.dtype),
                                   None, _lib.ptr(alpha), _lib.ptr(side), _lib.ptr(ga), _lib.ptr(gb),
                                   _lib.ptr(gp), _lib.dtype_enum(p.dtype), _lib.ptr(ge), _lib.dtype_enum(e.dtype),
                                   n, t, s, 1e-6, 0, flags, st)
    assert rc == 0


@pytest.mark.parametrize("s", [1024, 1000])
def test_training_calls_are_cuda_graph_capturable(s):
    """include/simulst_b200.h promises no host read and no hidden sync: capture fwd+bwd in a CUDA
    graph, replay it on NEW input values, compare with the eager call bit for bit."""
    from simulst_b200 import _lib
    lib = _lib.load()
    n, t = 6, 9
    flags = _lib.MMA_SOFT | _lib.MMA_MASS_PRESERVATION
    dt = torch.bfloat16

    def fresh(seed):
        p, se, _, ga, gb = _seeded(n, t, s, seed=seed)
        return p.to(DEV, dt), se.to(DEV, dt), ga.to(DEV), gb.to(DEV)

    p, e, ga, gb = fresh(1)
    alpha = torch.empty(n, t, s, device=DEV)
    beta = torch.empty_like(alpha)
    side = torch.empty(n, t, 2, device=DEV)
    gp, ge = torch.empty_like(p), torch.empty_like(e)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    bufs = (p, e, alpha, beta, side, ga, gb, gp, ge, status)
    side_stream = torch.cuda.Stream()
    with torch.cuda.stream(side_stream):
        _raw_train_call(lib, _lib, bufs, n, t, s, flags)          # warm-up (sets function attributes)
    side_stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        _raw_train_call(lib, _lib, bufs, n, t, s, flags)
    p2, e2, ga2, gb2 = fresh(2)
    p.copy_(p2); e.copy_(e2); ga.copy_(ga2); gb.copy_(gb2)
    for buf in (alpha, beta, gp, ge):
        buf.zero_()
    graph.replay()
    torch.cuda.synchronize()
    got = [x.clone() for x in (alpha, beta, gp, ge)]
    _raw_train_call(lib, _lib, bufs, n, t, s, flags)
    torch.cuda.synchronize()
    for a, b in zip(got, (alpha, beta, gp, ge)):
        assert torch.equal(a, b)
    assert int(status.item()) == 0


def test_step_call_is_cuda_graph_capturable():
    """One captured decoding step replayed 12 times carries head_step exactly like 12 eager calls."""
    from simulst_b200 import _lib
    lib = _lib.load()
    r, s = 64, 200
    g = torch.Generator().manual_seed(17)
    p_all = torch.sigmoid(torch.randn(12, r, s, generator=g) * 1.5 - 2.0).to(DEV)
    e_all = torch.randn(12, r, s, generator=g).to(DEV)
    p, e = p_all[0].clone(), e_all[0].clone()
    hs = torch.zeros(r, dtype=torch.long, device=DEV)
    hr = torch.empty(r, dtype=torch.uint8, device=DEV)
    alpha, beta = torch.empty_like(p), torch.empty_like(e)

    def call():
        rc = lib.simulst_mma_step(_lib.ptr(p), 0, _lib.ptr(e), 0, None, _lib.ptr(hs), _lib.ptr(hr), _lib.ptr(alpha),
                                  _lib.ptr(beta), r, s, _lib.MMA_MASS_PRESERVATION, _lib.stream_ptr(torch.device(DEV)))
        assert rc == 0

    side_stream = torch.cuda.Stream()
    with torch.cuda.stream(side_stream):
        call()
    side_stream.synchronize()
    hs.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call()
    hs_o = torch.zeros(r, dtype=torch.long)
    for k in range(12):
        p.copy_(p_all[k]); e.copy_(e_all[k])
        graph.replay()
        hs_o, hr_o, a_o, b_o = omma.mma_process_infer(p_all[k].cpu(), hs_o, e_all[k].cpu().unsqueeze(1), None, True)
        assert torch.equal(hs.cpu(), hs_o) and torch.equal(hr.cpu().bool(), hr_o)
        assert torch.equal(alpha.cpu(), a_o)
        torch.testing.assert_close(beta.cpu(), b_o.squeeze(1), rtol=1e-5, atol=1e-7)


# ----------------------------------------------------------------------------- left padding
@pytest.mark.parametrize("fused", [False, True])
def test_mass_preservation_left_padding(fused):
    """mass_preservation(alpha, padding_mask, left_padding=True) (monotonic_attention.py:183-185):
    padded columns are zeroed, the residual REPLACES the last column -- the no-mask rule, not the
    right-padding scatter_add."""
    from simulst_b200 import ops
    from simulst_b200.utils import monotonic_attention as ma
    n, t, s = 4, 6, 96
    g = torch.Generator().manual_seed(23)
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2.0)
    lens = torch.tensor([96, 70, 51, 96])
    mask = torch.arange(s)[None, :] < (s - lens)[:, None]          # padding on the LEFT
    ga = torch.randn(n, t, s, generator=g)
    p_o = p.clone().requires_grad_()
    a_o = omma.expected_alignment_from_p_choose(p_o, mask, eps=1e-6)
    a_o = omma.mass_preservation(a_o, mask, left_padding=True)
    (a_o * ga).sum().backward()
    p_d = p.to(DEV).requires_grad_()
    if fused:
        alpha, _ = ops.mma_train(p_d, None, mask.to(DEV), eps=1e-6, mass_preservation=True, left_padding=True)
    else:
        alpha = ma.expected_alignment_from_p_choose(p_d, mask.to(DEV), eps=1e-6)
        alpha = ma.mass_preservation(alpha, mask.to(DEV), left_padding=True)
    (alpha * ga.to(DEV)).sum().backward()
    assert_parity(alpha, a_o, "left-padding alpha")
    assert_parity(p_d.grad, p_o.grad, "left-padding grad_p", extra_atol=2e-6 * float(ga.abs().max()))
    # the same inputs went through the unmodified reference (tests/golden/mma_leftpad.npz)
    from tests.golden_io import load
    c = load("mma_leftpad.npz")["left"]
    assert torch.equal(c.p, p) and torch.equal(c.mask, mask)
    assert_parity(alpha, c.alpha, "left-padding alpha vs golden")
    assert_parity(p_d.grad, c.grad_p, "left-padding grad_p vs golden", extra_atol=2e-6 * float(ga.abs().max()))


def _random_dense_cases(count, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        s = int(torch.randint(1, 4300, (1,), generator=g))
        kind = int(torch.randint(0, 3, (1,), generator=g))
        if kind == 1:
            s = max(8, s // 8 * 8)          # rows the dense bulk-copy kernels take
        t = int(torch.randint(1, 5, (1,), generator=g))
        n = int(torch.randint(1, 4, (1,), generator=g))
        masked = bool(int(torch.randint(0, 2, (1,), generator=g)))
        dtype = [torch.float32, torch.bfloat16][int(torch.randint(0, 2, (1,), generator=g))]
        out.append((n, t, s, masked, dtype))
    return out


@pytest.mark.parametrize("n,t,s,masked,dtype", _random_dense_cases(40, 77), ids=lambda v: str(v).replace("torch.", ""))
def test_random_source_lengths_match_oracle(n, t, s, masked, dtype):
    """Seeded random source lengths 1..4300 -- every CTA size, rows that are and are not 16-byte multiples
    (bulk copies of the aligned superset), ragged tails, right-padded masks -- against the oracle."""
    p, se, _, ga, gb = _seeded(n, t, s, seed=9000 + s)
    p, se = p.to(dtype), se.to(dtype)
    mask = None
    if masked:
        g = torch.Generator().manual_seed(s)
        lens = torch.randint(1, s + 1, (n,), generator=g)
        lens[0] = s
        mask = torch.arange(s)[None, :] >= lens[:, None]
    a_o, b_o, gp_o, ge_o = _oracle(p.float(), se.float(), mask, ga, gb)
    a64, b64, gp64, ge64 = _oracle(p.float(), se.float(), mask, ga, gb, dtype64=True)
    alpha, beta, gp, ge = _run(p, se, mask, True, 0, True, ga, gb, dtype=dtype)
    tag = f"random dense n{n} T{t} S{s} m{int(masked)} {str(dtype)[6:]}"
    assert_parity(alpha, a_o, tag + " alpha", a64)
    assert_parity(beta, b_o, tag + " beta", b64)
    floor = grad_floor(s, ga, gb)
    rt = 1e-5 if dtype == torch.float32 else 2.0 ** -8
    assert_parity(gp, gp_o, tag + " grad_p", gp64, rtol=rt, extra_atol=floor)
    assert_parity(ge, ge_o, tag + " grad_energy", ge64, rtol=rt, extra_atol=floor)
