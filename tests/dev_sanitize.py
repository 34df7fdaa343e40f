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
simulst_b200.check_status(dev)
print("sanitize probe done")
