"""compute-sanitizer probe for the kernels of the end of round 2: SHIFT instantiations of the dense kernels
(unaligned rows, pitched outputs) and the masked rows that copy only their live bytes (neutral ring tails +
fixer stores).  Usage on a B200:  compute-sanitizer --tool memcheck|racecheck|synccheck python tests/dev_sanitize_shift.py"""
import sys
import torch
sys.path.insert(0, ".")
import simulst_b200
from simulst_b200 import ops

dev = torch.device("cuda")
g = torch.Generator().manual_seed(5)


def run(n, t, s, dtype, soft=True, mp=True, masked=False, delays=False, lens=None):
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2).to(dev, dtype).requires_grad_()
    e = torch.randn(n, t, s, generator=g).to(dev, dtype).requires_grad_() if soft else None
    mask = None
    if masked:
        L = lens if lens is not None else torch.randint(1, s + 1, (n,), generator=g)
        mask = (torch.arange(s)[None, :] >= L[:, None]).to(dev)
    if delays:
        a, b, d = ops.mma_train_with_delays(p, e, mask, mass_preservation=mp)
        loss = d.sum() + (b.sum() if soft else 0.0)
    else:
        a, b = ops.mma_train(p, e, mask, mass_preservation=mp)
        loss = (a * 0.5).sum() + (b.sum() if soft else 0.0)
    loss.backward()
    torch.cuda.synchronize()
    assert not torch.isnan(p.grad.float()).any()


for promise in (False, True):
    simulst_b200.assume_right_padding(promise)
    run(3, 4, 1500, torch.bfloat16, masked=promise)            # rows at 0 / 8 bytes, the row ends inside a thread
    run(3, 4, 1001, torch.bfloat16, masked=True)               # odd element offsets (byte permute)
    run(2, 4, 999, torch.float32, masked=promise)              # 4-byte offsets
    run(4, 3, 37, torch.float16, masked=True)                  # 8-byte units, one warp
    run(1, 2, 5003, torch.float32)                             # 12 elements per thread
    run(2, 3, 3001, torch.bfloat16, soft=False)                # hard attention: the fixer looks ahead alone
    run(2, 3, 1500, torch.bfloat16, delays=True, masked=True)
    # aligned rows with a mask: live-byte copies, fixer, rows without live columns next to full rows
    run(4, 5, 1024, torch.bfloat16, masked=True, lens=torch.tensor([1024, 771, 1, 512]))
    run(3, 4, 512, torch.float32, masked=True, lens=torch.tensor([0, 509, 512]))
    run(2, 3, 6000, torch.bfloat16, masked=True, lens=torch.tensor([6000, 4099]))
simulst_b200.assume_right_padding(False)
simulst_b200.check_status()
print("sanitize shift probe done")
