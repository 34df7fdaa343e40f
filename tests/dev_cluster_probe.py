"""Forward at SURVEY's long-form row count (64 rows) with and without the cluster kernel.  Not a test."""
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


def timeit(fn, reps=7):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]


for S in (1280, 2048, 3000, 4096, 6000, 8192):
    for T in (128,):
        g = torch.Generator().manual_seed(1)
        p = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, torch.bfloat16)
        e = torch.randn(N, T, S, generator=g).to(dev, torch.bfloat16)
        alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
        side = torch.empty(N, T, 2, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)

        def fwd():
            rc = lib.simulst_mma_train_fwd(p.data_ptr(), 1, e.data_ptr(), 1, None, alpha.data_ptr(), beta.data_ptr(),
                                           side.data_ptr(), N, T, S, 1e-6, 0, 3, status.data_ptr(), st)
            assert rc == 0, rc
        out = {"rows": N, "tgt": T, "src": S}
        for mode, name in ((0, "single_cta_us"), (2, "cluster_us")):
            lib.simulst_mma_set_cluster(mode)
            fwd(); torch.cuda.synchronize()
            out[name] = round(timeit(fwd), 1)
        out["speedup"] = round(out["single_cta_us"] / out["cluster_us"], 2)
        if len(sys.argv) > 2:           # every shape that holds the row
            for cl in (2, 4, 8):
                for th in (96, 128):
                    if cl * th * 8 < S or (cl == 8 and th > 128) or cl * th * 8 >= 2 * S + 2048:
                        continue
                    assert lib.simulst_mma_set_cluster_shape(cl, th) == 0
                    fwd(); torch.cuda.synchronize()
                    out[f"cl{cl}x{th}"] = round(timeit(fwd), 1)
            lib.simulst_mma_set_cluster_shape(0, 0)
        print(json.dumps(out), flush=True)
lib.simulst_mma_set_cluster(1)
