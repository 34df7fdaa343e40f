"""a8-a10 on the REAL reference classes: the unmodified ``MonotonicAttention`` /
``MonotonicInfiniteLookbackAttention`` (modules/monotonic_multihead_attention.py:29,460) with the
two method bodies replaced by the sm_100a path, driven through the reference's own ``forward()``
(:354-423) on the GPU, against the untouched class running on the CPU -- training pass with
gradients of every parameter, and a run of incremental decoding steps carrying ``head_step`` /
``head_read`` in ``incremental_state``.

The reference files are executed from ``/root/reference`` in the build container and from
``baseline/_ref`` (git-ignored copy made by ``oracle/ship_reference.py``) on the GPU box."""
import copy

import pytest
import torch

from oracle import ref_loader
from tests.parity import assert_parity

pytestmark = [pytest.mark.gpu, pytest.mark.reference]
DEV = "cuda"


@pytest.fixture(autouse=True)
def _ieee_fp32_around_the_hot_path():
    """The projections / convolutions AROUND the hot path run in cuBLAS / cuDNN here and in MKL on
    the CPU side; TF32 (cuDNN's default for convolutions) would put 1e-3 relative noise into the
    inputs of the path under test."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _pair(kind, how, mass_preservation=True, heads=4, embed=64, seed=3):
    """(reference module on CPU, same weights on the GPU with the kernel-backed method bodies)."""
    from simulst_b200.modules.monotonic_multihead_attention import (
        B200MonotonicAttentionMixin, patch_monotonic_attention)
    ref = ref_loader.make_attention(kind, embed, heads, mass_preservation=mass_preservation, seed=seed)
    ref.noise_std = 0.0          # training-mode noise is drawn on the module's device: CPU and CUDA streams differ
    mine = copy.deepcopy(ref)
    base = type(ref)
    if how == "mixin":
        cls = type("B200" + base.__name__, (B200MonotonicAttentionMixin, base), {})
    else:
        cls = patch_monotonic_attention(type("Patched" + base.__name__, (base,), {}))
    mine.__class__ = cls
    return ref, mine.to(DEV)


def _inputs(t, s, bsz, embed, seed, masked):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(t, bsz, embed, generator=g)
    k = torch.randn(s, bsz, embed, generator=g)
    mask = None
    if masked:
        lens = torch.randint(max(2, s // 2), s + 1, (bsz,), generator=g)
        lens[0] = s
        mask = torch.arange(s)[None, :] >= lens[:, None]
    return q, k, mask


@pytest.mark.parametrize("kind", ["infinite_lookback", "hard_aligned"])
@pytest.mark.parametrize("how", ["mixin", "patch"])
@pytest.mark.parametrize("masked", [False, True])
def test_reference_forward_training_pass(kind, how, masked):
    ref, mine = _pair(kind, how)
    ref.train(); mine.train()
    t, s, bsz, embed = 11, 72, 3, 64
    q, k, mask = _inputs(t, s, bsz, embed, 5, masked)
    g = torch.Generator().manual_seed(6)
    q_r, k_r = q.clone().requires_grad_(), k.clone().requires_grad_()
    out_r, extra_r = ref(q_r, k_r, k_r, key_padding_mask=mask)
    w_out = torch.randn(out_r.shape, generator=g)
    w_alpha = torch.randn(extra_r["alpha"].shape, generator=g) * 0.1
    (out_r * w_out).sum().add((extra_r["alpha"] * w_alpha).sum()).backward()

    q_m, k_m = q.to(DEV).requires_grad_(), k.to(DEV).requires_grad_()
    out_m, extra_m = mine(q_m, k_m, k_m, key_padding_mask=mask.to(DEV) if masked else None)
    (out_m * w_out.to(DEV)).sum().add((extra_m["alpha"] * w_alpha.to(DEV)).sum()).backward()

    # the projections and the three bmm's run on different back-ends (CPU MKL vs cuBLAS TF32-off fp32):
    # their own rounding enters at ~1e-6 relative to the activations, so the gate keeps rtol 1e-5 with
    # an absolute floor of 5e-6 x scale here
    kw = dict(atol=5e-6)
    for key in ("p_choose", "alpha", "beta"):
        assert extra_m[key].shape == extra_r[key].shape
        assert_parity(extra_m[key], extra_r[key].detach(), f"{kind}/{how} {key}", **kw)
    assert_parity(out_m, out_r.detach(), f"{kind}/{how} attn", **kw)
    assert_parity(q_m.grad, q_r.grad, "grad query", rtol=1e-4, atol=2e-5)
    assert_parity(k_m.grad, k_r.grad, "grad key", rtol=1e-4, atol=2e-5)
    # some parameter gradients are analytically zero (a key bias shifts every energy of a query
    # alike): both sides hold rounding noise there, so the floor is tied to the largest gradient
    g_scale = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    for (name, p_r), (_, p_m) in zip(ref.named_parameters(), mine.named_parameters()):
        if p_r.grad is None:
            assert p_m.grad is None, name
            continue
        assert_parity(p_m.grad, p_r.grad, f"grad {name}", rtol=1e-4, atol=2e-5, extra_atol=2e-6 * g_scale)


@pytest.mark.parametrize("kind", ["infinite_lookback", "hard_aligned"])
@pytest.mark.parametrize("mass_preservation", [True, False])
def test_reference_forward_incremental_steps(kind, mass_preservation):
    """24 decoding steps through the reference's forward(..., incremental_state=...) on a growing
    source prefix (what the SimulEval agent does): the `head_step` / `head_read` caches, the one-hot
    alpha and the attention output follow the untouched class step by step."""
    ref, mine = _pair(kind, "mixin", mass_preservation=mass_preservation)
    ref.eval(); mine.eval()
    bsz, embed, s_max, steps = 5, 64, 60, 24
    g = torch.Generator().manual_seed(9)
    keys = torch.randn(s_max, bsz, embed, generator=g) * 2.0
    queries = torch.randn(steps, 1, bsz, embed, generator=g) * 2.0
    lens = torch.randint(s_max // 2, s_max + 1, (bsz,), generator=g)
    lens[0] = s_max
    mask = torch.arange(s_max)[None, :] >= lens[:, None]
    inc_r, inc_m = {}, {}
    with torch.no_grad():
        for st in range(steps):
            out_r, ex_r = ref(queries[st], keys, keys, key_padding_mask=mask, incremental_state=inc_r)
            out_m, ex_m = mine(queries[st].to(DEV), keys.to(DEV), keys.to(DEV), key_padding_mask=mask.to(DEV),
                               incremental_state=inc_m)
            buf_r, buf_m = ref._get_monotonic_buffer(inc_r), mine._get_monotonic_buffer(inc_m)
            assert torch.equal(buf_m["head_step"].cpu(), buf_r["head_step"]), st
            assert torch.equal(buf_m["head_read"].cpu(), buf_r["head_read"]), st
            assert buf_m["head_step"].shape == buf_r["head_step"].shape == (bsz, ref.num_heads)
            assert torch.equal(ex_m["alpha"].cpu(), ex_r["alpha"]), st
            assert_parity(ex_m["beta"], ex_r["beta"], f"step {st} beta", atol=5e-6)
            assert_parity(out_m, out_r, f"step {st} attn", atol=5e-6)


def test_reference_forward_with_expected_delays_epilogue():
    """Opt-in on the patched reference class: the kernel leaves sum_j (j+1)*alpha on the module."""
    ref, mine = _pair("infinite_lookback", "patch")
    ref.train(); mine.train()
    mine.with_expected_delays = True
    q, k, mask = _inputs(7, 40, 2, 64, 12, True)
    _, ex_r = ref(q, k, k, key_padding_mask=mask)
    _, ex_m = mine(q.to(DEV), k.to(DEV), k.to(DEV), key_padding_mask=mask.to(DEV))
    steps = torch.arange(1, 41).float()
    want = (ex_r["alpha"].detach() * steps).sum(-1).view(-1, 7)
    assert_parity(mine.expected_delays, want, "expected_delays", atol=5e-6)


# ----------------------------------------------------------------------------- CIFLayer (a13 / a14)
def _cif_pair(sg_alpha=False, beta=1.0, c=32, seed=4):
    """(reference CIFLayer on CPU, same weights on the GPU behind B200CIFLayerMixin)."""
    from simulst_b200.models.cif_transformer import B200CIFLayerMixin
    CIFLayer, _ = ref_loader.load_cif_layer()
    torch.manual_seed(seed)
    ref = CIFLayer(c, c, 3, 0.0, sg_alpha, beta)
    with torch.no_grad():
        ref.alpha_proj[-1].bias.fill_(-0.5)
    mine = copy.deepcopy(ref)
    mine.__class__ = type("B200CIFLayer", (B200CIFLayerMixin, CIFLayer), {})
    return ref, mine.to(DEV)


@pytest.mark.parametrize("sg_alpha", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_cif_layer_mixin_on_reference_class_forward(sg_alpha, masked):
    ref, mine = _cif_pair(sg_alpha)
    ref.train(); mine.train()
    s, b, c = 90, 4, 32
    g = torch.Generator().manual_seed(14)
    x = torch.randn(s, b, c, generator=g)
    mask = None
    if masked:
        lens = torch.tensor([90, 61, 77, 50])
        mask = torch.arange(s)[None, :] >= lens[:, None]
    x_r = x.clone().requires_grad_()
    with torch.no_grad():
        a = ref.alpha_proj(x).transpose(1, 0).sigmoid().squeeze(-1)
        if masked:
            a = a.masked_fill(mask, 0)
        tl = a.sum(1).round().clamp(min=1).long()
    out_r = ref(x_r, mask, tl)
    w = torch.randn(out_r["cif_out"][0].shape, generator=g)
    (out_r["cif_out"][0] * w).sum().backward()
    x_m = x.to(DEV).requires_grad_()
    out_m = mine(x_m, mask.to(DEV) if masked else None, tl.to(DEV))
    (out_m["cif_out"][0] * w.to(DEV)).sum().backward()
    assert torch.equal(out_m["cif_lengths"][0].cpu(), out_r["cif_lengths"][0])
    t_len = int(tl.max())
    slack = 8 * 2.0 ** -23 * t_len * float(x.abs().max())       # 4 ulp of the running sum T*beta, both sides
    assert_parity(out_m["cif_out"][0], out_r["cif_out"][0].detach(), "cif_out", extra_atol=slack, atol=5e-6)
    assert_parity(out_m["alpha"][0], out_r["alpha"][0].detach(), "alpha", atol=5e-6)
    assert_parity(x_m.grad, x_r.grad, "grad x", rtol=1e-4, atol=1e-4)
    for (name, p_r), (_, p_m) in zip(ref.named_parameters(), mine.named_parameters()):
        # (sg_alpha only detaches x on its way INTO alpha_proj: the projection's own parameters
        # still receive the CIF gradient through alpha)
        assert_parity(p_m.grad, p_r.grad, f"grad {name}", rtol=1e-4, atol=1e-4)


def test_cif_layer_mixin_on_reference_class_streaming():
    """6 chunks through ``infer`` with the carry in incremental_state, then ``finish``; calling
    again after finish raises in both (reference :254 leaves None in the cache)."""
    ref, mine = _cif_pair()
    ref.eval(); mine.eval()
    g = torch.Generator().manual_seed(15)
    x = torch.randn(96, 1, 32, generator=g)
    inc_r, inc_m = {}, {}
    with torch.no_grad():
        for k in range(6):
            ch = x[k * 16:(k + 1) * 16]
            o_r = ref.infer(ch, inc_r, None, finish=(k == 5))
            o_m = mine.infer(ch.to(DEV), inc_m, None, finish=(k == 5))
            assert torch.equal(o_m["cif_lengths"][0].cpu(), o_r["cif_lengths"][0]), k
            assert o_m["cif_out"][0].shape == o_r["cif_out"][0].shape
            assert_parity(o_m["cif_out"][0], o_r["cif_out"][0], f"chunk {k}", atol=1e-5)
        with pytest.raises(AttributeError):
            ref.infer(x[:16], inc_r, None)
        with pytest.raises(AttributeError):
            mine.infer(x[:16].to(DEV), inc_m, None)


def test_cif_infer_zero_tail_weight_quirk():
    """tail_weight == 0 with tail_thres == 0 gives beta/0 -> 0*inf = NaN in the
    carried feature.  The decision is REPRODUCE: reference, oracle and kernel agree on the NaNs."""
    from oracle import cif as ocif
    from simulst_b200.models.torch_cif import cif_function
    rc = ref_loader.load_cif()
    x = torch.ones(1, 4, 3)
    alpha = torch.tensor([[0.5, 0.5, 0.5, 0.5]])        # fires exactly at frames 1 and 3: nothing left over
    want = rc.cif_function(x, alpha, beta=1.0, tail_thres=0.0)
    ora = ocif.cif_function(x, alpha, beta=1.0, tail_thres=0.0)
    got = cif_function(x.to(DEV), alpha.to(DEV), beta=1.0, tail_thres=0.0)
    assert float(want["tail_weights"][0]) == 0.0
    assert bool(torch.isnan(want["cif_out"][0][0, -1]).all())
    for res in (ora, got):
        assert torch.equal(res["cif_lengths"][0].cpu(), want["cif_lengths"][0])
        assert torch.equal(torch.isnan(res["cif_out"][0]).cpu(), torch.isnan(want["cif_out"][0]))
        torch.testing.assert_close(res["cif_out"][0].cpu()[:, :-1], want["cif_out"][0][:, :-1])
