"""Small batches: does a row finish its steps faster on twice the threads with 4 elements each?  Not a test."""
import json
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 128


def timeit(fn, reps=7):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]


lib.simulst_mma_set_cluster(0)
for S, cfg in ((512, None), (1024, (256, 4)), (2048, (512, 4))):
    g = torch.Generator().manual_seed(1)
    p = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, torch.bfloat16)
    e = torch.randn(N, T, S, generator=g).to(dev, torch.bfloat16)
    alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
    side = torch.empty(N, T, 2, device=dev)
    ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
    gp = torch.empty_like(p); ge = torch.empty_like(e)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def fwd():
        assert lib.simulst_mma_train_fwd(p.data_ptr(), 1, e.data_ptr(), 1, None, alpha.data_ptr(), beta.data_ptr(),
                                         side.data_ptr(), N, T, S, 1e-6, 0, 3, status.data_ptr(), st) == 0

    def bwd():
        assert lib.simulst_mma_train_bwd(p.data_ptr(), 1, e.data_ptr(), 1, None, alpha.data_ptr(), side.data_ptr(),
                                         ga.data_ptr(), gb.data_ptr(), gp.data_ptr(), 1, ge.data_ptr(), 1,
                                         N, T, S, 1e-6, 0, 3, st) == 0
    out = {"rows": N, "src": S}
    for name, c in (("default", None), ("vpt4", cfg)):
        if name == "vpt4" and c is None:
            continue
        assert lib.simulst_mma_set_config(*(c or (0, 0))) == 0
        fwd(); bwd(); torch.cuda.synchronize()
        out[name] = {"cfg": c, "fwd_us": round(timeit(fwd), 1), "bwd_us": round(timeit(bwd), 1)}
    lib.simulst_mma_set_config(0, 0)
    print(json.dumps(out), flush=True)
lib.simulst_mma_set_cluster(1)
