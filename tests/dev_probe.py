"""Development probe (not a test): run on the GPU box to print error magnitudes."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import mma as omma
from simulst_b200 import ops, _lib
from tests.test_mma_train_gpu import _seeded, _run

def report(tag, got, ref, ref64=None):
    d = (got.double() - ref.double()).abs()
    rel = d / ref.double().abs().clamp_min(1e-30)
    big = ref.abs() > 1e-3
    msg = f"{tag}: max_abs={d.max().item():.3e} max_rel(|ref|>1e-3)={rel[big].max().item() if big.any() else 0:.3e}"
    if ref64 is not None:
        msg += f" | vs64 kernel={(got.double()-ref64).abs().max().item():.3e} ref={(ref.double()-ref64).abs().max().item():.3e}"
    print(msg, flush=True)

for (n, t, s, masked, chunk) in [(8, 32, 256, False, 0), (4, 128, 1024, False, 0), (4, 64, 1024, True, 0), (2, 16, 300, True, 5)]:
    p, se, mask, ga, gb = _seeded(n, t, s, seed=1234, masked=masked)
    alpha, beta, gp, ge = _run(p, se, mask, True, chunk, True, ga, gb)
    p_o = p.clone().requires_grad_(); se_o = se.clone().requires_grad_()
    a_o, b_o = omma.mma_process_train(p_o, se_o, mask, 1e-6, True, chunk or None)
    ((a_o * ga).sum() + (b_o * gb).sum()).backward()
    p6 = p.double().requires_grad_(); s6 = se.double().requires_grad_()
    a64, b64 = omma.mma_process_train(p6, s6, mask, 1e-6, True, chunk or None, compute_dtype=torch.float64)
    ((a64 * ga).sum() + (b64 * gb).sum()).backward()
    print(f"--- N={n} T={t} S={s} masked={masked} chunk={chunk}")
    report("alpha", alpha, a_o.detach(), a64.detach())
    report("beta ", beta, b_o.detach(), b64.detach())
    report("gp   ", gp, p_o.grad, p6.grad)
    report("ge   ", ge, se_o.grad, s6.grad)
