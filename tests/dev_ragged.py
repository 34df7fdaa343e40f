"""Development probe: forward pipelined kernel on ragged dense rows vs the generic kernel."""
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda")
for (n, t, s, soft) in [(2, 12, 96, True), (3, 7, 1000, False), (3, 7, 1000, True), (2, 5, 1504, True), (2, 3, 6000, True)]:
    g = torch.Generator().manual_seed(5)
    p = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2).to(dev)
    e = torch.randn(n, t, s, generator=g).to(dev) if soft else None
    outs = []
    for mode in (1, 0):
        lib.simulst_mma_set_pipeline(mode)
        a, b, d = ops.mma_train_with_delays(p, e, None)
        torch.cuda.synchronize()
        outs.append((a.clone(), b.clone(), d.clone()))
    lib.simulst_mma_set_pipeline(5)
    for name, x, y in zip(("alpha", "beta", "delays"), outs[0], outs[1]):
        nan = int(torch.isnan(x).sum())
        diff = float((x - y).abs().max()) if nan == 0 else float("nan")
        where = ""
        if nan:
            idx = torch.nonzero(torch.isnan(x))[:3].tolist()
            where = f" first NaN at {idx}"
        print(f"({n},{t},{s},soft={soft}) {name}: NaNs {nan} max|pipe-generic| {diff:.3e}{where}", flush=True)
