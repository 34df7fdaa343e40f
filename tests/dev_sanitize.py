"""Development probe for compute-sanitizer (memcheck / racecheck / synccheck): one small call of
every kernel family added this round.  Usage on a B200:
    compute-sanitizer --tool racecheck python tests/dev_sanitize.py"""
import sys
import torch
sys.path.insert(0, ".")
import simulst_b200
from simulst_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(3)


def run(n, t, s, soft, mp, masked, dtype=torch.float32, holes=False, delays=False):
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2).to(dev, dtype).requires_grad_()
    e = torch.randn(n, t, s, generator=g).to(dev, dtype).requires_grad_() if soft else None
    mask = None
    if masked:
        lens = torch.randint(s // 2, s + 1, (n,), generator=g)
        mask = torch.arange(s)[None, :] >= lens[:, None]
        if holes:
            mask[-1, 3] = True
        mask = mask.to(dev)
    if delays:
        a, b, d = ops.mma_train_with_delays(p, e, mask, mass_preservation=mp)
        loss = d.sum() + (b.sum() if soft else 0.0)
    else:
        a, b = ops.mma_train(p, e, mask, mass_preservation=mp)
        loss = (a * 0.5).sum() + (b.sum() if soft else 0.0)
    loss.backward()
    torch.cuda.synchronize()
    assert not torch.isnan(p.grad.float()).any()


for mode in (5, 0):
    lib.simulst_mma_set_pipeline(mode)
    run(2, 3, 256, True, True, False)                       # dense, 1 warp
    run(2, 3, 1024, True, True, False, torch.bfloat16)      # dense, 4 warps
    run(2, 2, 264, True, True, False)                       # ragged
    run(1, 2, 4096, True, True, False, torch.bfloat16)      # 16 warps
    run(1, 2, 6000, False, True, False, torch.bfloat16)     # 12 elements per thread, 2-stage ring
    run(3, 3, 512, True, True, True, holes=True)            # masked: split by row
    run(2, 3, 512, True, False, False, delays=True)         # delay epilogue
lib.simulst_mma_set_pipeline(5)

# ---- round 2: CTA sizes in one-warp steps, rows that are not 16-byte multiples (shifted bulk copies),
# the pooled-grid kernels (warp-specialised / single-warp / generic recurrence kernels, lean and generic
# row kernels, right-padded rows with the residual on and off the grid), latency / SSNT / CTC kernels
run(2, 3, 768, True, True, False, torch.bfloat16)       # 96 threads
run(2, 2, 1504, True, True, False)                      # 192 threads
run(1, 2, 3000, True, True, False)                      # 384 threads, ragged
run(2, 3, 999, True, True, False)                       # rows not 16-byte multiples
run(2, 3, 1500, True, True, True, torch.bfloat16)


def run_pooled(n, t, s, ratio, soft, masked, dtype=torch.float32, lean=False):
    sp = (s + ratio - 1) // ratio
    pp = torch.sigmoid(torch.randn(n, t, sp, generator=g) - 1).to(dev, dtype).requires_grad_()
    e = torch.randn(n, t, s, generator=g).to(dev, dtype).requires_grad_() if soft else None
    mask = None
    if masked:
        lens = torch.randint(s // 2, s + 1, (n,), generator=g)
        mask = (torch.arange(s)[None, :] >= lens[:, None]).to(dev)
    _, a, b, d = ops.mma_train_pooled(pp, s, ratio, e, mask, with_delays=lean, want_dense=not lean,
                                      right_padding=masked, want_alpha=not lean)
    loss = (d.sum() if lean else (a * 0.5).sum()) + (b.sum() if soft else 0.0)
    loss.backward()
    torch.cuda.synchronize()
    assert not torch.isnan(pp.grad.float()).any()


run_pooled(2, 9, 1024, 8, True, False, torch.bfloat16)      # warp-specialised K1/K4 + lean row kernels
run_pooled(2, 9, 1024, 8, True, False, torch.bfloat16, lean=True)
run_pooled(3, 5, 1000, 8, True, True)                       # right padding, residual off the grid
run_pooled(2, 5, 2048, 8, True, False)                      # Sp = 256: 8 elements per lane
run_pooled(2, 4, 1024, 3, True, True)                       # generic row kernels
run_pooled(2, 4, 4096, 8, True, False)                      # Sp = 512: generic multi-warp recurrence kernels
run_pooled(2, 5, 520, 16, False, True)                      # hard-aligned

from simulst_b200.criterion.best_alignment import best_alignment
from simulst_b200.criterion.ssnt_loss import ssnt_loss
lp = torch.randn(40, 3, 32, generator=g).log_softmax(-1).to(dev)
best_alignment(lp, torch.randint(1, 32, (3, 7), generator=g).to(dev), torch.tensor([40, 33, 21]).to(dev),
               torch.tensor([7, 5, 3]).to(dev))
lpp = torch.randn(2, 5, 24, 16, generator=g).to(dev).log_softmax(-1).requires_grad_()
em = torch.randn(2, 5, 24, generator=g).to(dev).requires_grad_()
loss, _, _ = ssnt_loss(lpp, torch.randint(0, 16, (2, 5), generator=g).to(dev), torch.tensor([24, 17]).to(dev),
                       torch.tensor([5, 3]).to(dev), emit_logits=em, reduction="sum")
loss.backward()
dl = torch.rand(6, 12, generator=g).cumsum(1).to(dev).requires_grad_()
ops.differentiable_average_lagging(dl, torch.full((6,), 40).to(dev)).sum().backward()
torch.cuda.synchronize()
simulst_b200.check_status(dev)
print("sanitize probe done")
