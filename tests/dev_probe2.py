"""Development probe (not a test): error magnitudes of the failing cases vs fp32 and fp64 oracles."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import mma as omma, cif as ocif
from tests.golden_io import load, opt
from tests.test_mma_train_gpu import _seeded, _run
from tests.test_cif_gpu import _seeded as cseed, _run as crun

def rep(tag, got, ref, r64=None):
    d = (got.double() - ref.double()).abs()
    msg = f"  {tag}: scale={ref.abs().max().item():.3e} max_abs={d.max().item():.3e} n_bad(1e-5 scale)={(d > 1e-5*ref.abs().max()+1e-5*ref.abs()).sum().item()}"
    if r64 is not None:
        msg += f" | vs64: kernel={(got.double()-r64).abs().max().item():.3e} ref={(ref.double()-r64).abs().max().item():.3e}"
    print(msg, flush=True)

print("== MMA S=6000 masked")
p, se, mask, ga, gb = _seeded(2, 4, 6000, seed=100 + 6000 + 4, masked=True)
alpha, beta, gp, ge = _run(p, se, mask, True, 0, True, ga, gb)
po = p.clone().requires_grad_(); so = se.clone().requires_grad_()
a_o, b_o = omma.mma_process_train(po, so, mask, 1e-6, True, None); ((a_o*ga).sum()+(b_o*gb).sum()).backward()
p6 = p.double().requires_grad_(); s6 = se.double().requires_grad_()
a6, b6 = omma.mma_process_train(p6, s6, mask, 1e-6, True, None, compute_dtype=torch.float64); ((a6*ga).sum()+(b6*gb).sum()).backward()
rep("alpha", alpha, a_o.detach(), a6.detach()); rep("beta", beta, b_o.detach(), b6.detach())
rep("gp", gp, po.grad, p6.grad); rep("ge", ge, so.grad, s6.grad)

print("== CIF")
for (b, s, c, beta, train) in [(8, 1500, 256, 1.0, True), (4, 700, 80, 0.35, True), (3, 333, 7, 1.3, True), (8, 1500, 256, 1.0, False)]:
    x, a, mask, g = cseed(b, s, c, seed=2024 + s + c)
    tl = (a.masked_fill(mask, 0).sum(1) / beta).round().clamp(min=1).long() if train else None
    xo = x.clone().requires_grad_(); ao = a.clone().requires_grad_()
    ref = ocif.cif_function(xo, ao, beta=beta, tail_thres=beta/2, padding_mask=mask, target_lengths=tl)
    g_out = torch.randn(ref["cif_out"][0].shape, generator=g); g_delay = torch.randn(ref["delays"][0].shape, generator=g)
    ((ref["cif_out"][0]*g_out).sum()+(ref["delays"][0]*g_delay).sum()).backward()
    x6 = x.double().requires_grad_(); a6 = a.double().requires_grad_()
    r6 = ocif.cif_function(x6, a6, beta=beta, tail_thres=beta/2, padding_mask=mask, target_lengths=tl, compute_dtype=torch.float64)
    ((r6["cif_out"][0]*g_out).sum()+(r6["delays"][0]*g_delay).sum()).backward()
    res, (gx, gal) = crun(x, a, beta, beta/2, mask, tl, g_out, g_delay)
    print(f"-- B={b} S={s} C={c} beta={beta} train={train} T={ref['cif_out'][0].shape[1]} lens_equal={torch.equal(res['cif_lengths'][0].cpu(), ref['cif_lengths'][0])} shape64_equal={r6['cif_out'][0].shape==ref['cif_out'][0].shape}")
    same = r6['cif_out'][0].shape==ref['cif_out'][0].shape
    rep("cif_out", res["cif_out"][0].detach().cpu(), ref["cif_out"][0].detach(), r6["cif_out"][0].detach() if same else None)
    rep("delays", res["delays"][0].detach().cpu(), ref["delays"][0].detach(), r6["delays"][0].detach() if same else None)
    rep("gx", gx, xo.grad, x6.grad if same else None); rep("ga", gal, ao.grad, a6.grad if same else None)
