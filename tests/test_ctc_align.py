"""CTC best alignment.

CPU: the oracle's wrapper restatement against the reference's OWN wrapper source
(best_alignment/__init__.py:25-111, executed over the oracle's forward pass), and the oracle's
Viterbi against exhaustive search.  GPU: the kernel against the oracle (bit-exact states), and --
when baseline/_ref travelled and nvcc can JIT it -- against the reference's own CUDA kernel + Python."""
import pytest
import torch

from oracle import ctc_align as oc
from oracle import ref_loader


def _case(g, n, s, v, t_max, ragged=True, dup=False):
    lp = (torch.randn(s, n, v, generator=g) * 2).log_softmax(-1)
    tl = torch.randint(0 if ragged else t_max, t_max + 1, (n,), generator=g)
    il = torch.randint(max(1, s // 2) if ragged else s, s + 1, (n,), generator=g)
    tl[0], il[0] = t_max, s
    tg = torch.randint(1, v, (n, t_max), generator=g)
    if dup and t_max > 1:
        tg[:, 1::2] = tg[:, 0:-1:2][:, :tg[:, 1::2].shape[1]]        # repeated labels: the +2 jump is illegal
    return lp, tg, il, tl


@pytest.mark.reference
def test_oracle_wrapper_matches_reference_wrapper_source():
    ref_py = ref_loader.load_best_alignment_python()
    g = torch.Generator().manual_seed(0)
    compared_labels = 0
    for trial in range(40):
        n, s, v, t_max = [int(torch.randint(lo, hi, (1,), generator=g)) for lo, hi in ((1, 5), (1, 14), (2, 6), (1, 6))]
        lp, tg, il, tl = _case(g, n, s, v, t_max, dup=trial % 3 == 0)
        assert torch.equal(oc.best_alignment(lp, tg, il, tl, 0), ref_py(oc.as_extension(), lp, tg, il, tl, 0, False))
        try:
            want = ref_py(oc.as_extension(), lp, tg, il, tl, 0, True)
        except RuntimeError:        # upstream: targets.gather one past the end (see simulst_b200/criterion/best_alignment.py)
            continue
        compared_labels += 1
        assert torch.equal(oc.best_alignment(lp, tg, il, tl, 0, True), want)
    assert compared_labels > 5


def test_oracle_viterbi_is_the_exhaustive_optimum():
    g = torch.Generator().manual_seed(1)
    checked = 0
    for trial in range(40):
        s = int(torch.randint(2, 7, (1,), generator=g))
        t = int(torch.randint(1, 4, (1,), generator=g))
        lp = torch.randn(s, 1, 4, generator=g).log_softmax(-1)
        tg = torch.randint(1, 4, (1, t), generator=g)
        best, path = oc.bruteforce_best_path(lp[:, 0], tg[0].tolist(), 0)
        if path is None:
            continue
        st = oc.best_alignment(lp, tg, torch.tensor([s]), torch.tensor([t]), 0)[0].tolist()
        aug = oc._augmented(tg[0].numpy(), t, 0)
        score = sum(float(lp[i, 0, aug[x]]) for i, x in enumerate(st))
        assert abs(score - best) < 1e-5, (st, path)
        checked += 1
    assert checked > 10


SHAPES = [
    # n, S, V, Tmax, ragged, dup
    (4, 13, 6, 3, True, False),
    (3, 50, 11, 9, True, True),
    (5, 200, 40, 30, True, False),
    (2, 64, 9, 1, False, False),
    (2, 7, 5, 4, True, False),         # more states than some samples have frames: partial alignment
    (3, 300, 50, 150, True, True),     # 301 states
    (1, 1, 3, 1, False, False),
    (2, 1500, 64, 200, True, False),   # BASELINE config 3's source length
]


@pytest.mark.gpu
@pytest.mark.parametrize("n,s,v,t_max,ragged,dup", SHAPES)
def test_kernel_states_and_labels_match_oracle(n, s, v, t_max, ragged, dup):
    from simulst_b200.criterion.best_alignment import best_alignment
    g = torch.Generator().manual_seed(100 + s + t_max)
    lp, tg, il, tl = _case(g, n, s, v, t_max, ragged, dup)
    want = oc.best_alignment(lp, tg, il, tl, 0)
    got = best_alignment(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0)
    assert got.dtype == torch.int64 and tuple(got.shape) == (n, s)
    assert torch.equal(got.cpu(), want)
    want_l = oc.best_alignment(lp, tg, il, tl, 0, True)
    got_l = best_alignment(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0, as_labels=True)
    assert torch.equal(got_l.cpu(), want_l)


@pytest.mark.gpu
def test_kernel_nll_and_quantity_targets():
    from simulst_b200 import ops
    from simulst_b200.criterion.best_alignment import quantity_targets
    g = torch.Generator().manual_seed(5)
    lp, tg, il, tl = _case(g, 4, 80, 12, 10)
    nll_o, _, _ = oc.viterbi_forward(lp, tg, il, tl, 0)
    states, nll = ops.ctc_best_alignment(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0, return_nll=True)
    torch.testing.assert_close(nll.cpu(), nll_o, rtol=1e-5, atol=1e-5)
    # cif_criterion.py:249-262 on the kernel's states: one boundary per aligned target token
    mask = torch.arange(80)[None, :] >= il[:, None]
    boundary, qt = quantity_targets(states, mask.cuda())
    seg = states.cpu().div(2, rounding_mode="floor")
    want_b = (seg != seg.roll(-1, dims=1)) & (states.cpu() % 2 != 0)
    want_b[mask] = False
    assert torch.equal(boundary.cpu(), want_b) and torch.equal(qt.cpu(), want_b.cumsum(1))
    assert bool((boundary.sum(1).cpu() <= tl).all())


@pytest.mark.gpu
def test_kernel_bf16_log_probs():
    from simulst_b200.criterion.best_alignment import best_alignment
    g = torch.Generator().manual_seed(6)
    lp, tg, il, tl = _case(g, 3, 120, 20, 12)
    lp16 = lp.to(torch.bfloat16)
    want = oc.best_alignment(lp16.float(), tg, il, tl, 0)
    got = best_alignment(lp16.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0)
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
@pytest.mark.reference
def test_kernel_matches_the_reference_cuda_kernel_and_python():
    """The real thing: the reference's best_alignment.cu/.cpp JIT-built with torch's cpp_extension
    (as best_alignment/__init__.py:10-17 does) and its own Python wrapper, on the same inputs."""
    import os
    if os.environ.get("SIMULST_SKIP_REF_JIT"):
        pytest.skip("reference JIT build disabled")
    try:
        ext = ref_loader.load_best_alignment_extension()
    except Exception as exc:      # no nvcc / ninja / headers on this box: cannot build the reference
        pytest.skip(f"cannot JIT-build the reference kernel here: {type(exc).__name__}: {str(exc)[:200]}")
    from simulst_b200.criterion.best_alignment import best_alignment
    ref_py = ref_loader.load_best_alignment_python()
    g = torch.Generator().manual_seed(7)
    for n, s, v, t_max, ragged, dup in SHAPES[:6]:
        lp, tg, il, tl = _case(g, n, s, v, t_max, ragged, dup)
        want = ref_py(ext, lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0, False)
        got = best_alignment(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0)
        assert torch.equal(got, want), (n, s, v, t_max)
        # and the oracle's forward pass against the reference kernel's outputs
        nll_r, la_r, paths_r = ext.best_alignment(lp.cuda(), tg.cuda(), il.cuda(), tl.cuda(), 0, True)
        nll_o, la_o, paths_o = oc.viterbi_forward(lp, tg, il, tl, 0)
        torch.testing.assert_close(nll_r.cpu(), nll_o, rtol=1e-5, atol=1e-5)
        assert torch.equal(paths_r.cpu(), paths_o)
        torch.testing.assert_close(la_r.cpu(), la_o, rtol=1e-6, atol=1e-6)
