"""Masked (right-padded) training shape, fwd and bwd timed separately through the dense C-ABI entry points,
for any build of the library (SIMULST_LIB=path).  Lengths: full, U[S/2,S], fixed 3S/4.  Not a test."""
import ctypes
import json
import os
import sys
import torch

path = os.environ.get("SIMULST_LIB", "simulst_b200/libsimulst_b200.so")
lib = ctypes.CDLL(os.path.abspath(path))
vp, ci, cf, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint
lib.simulst_mma_train_fwd.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, ci, ci, ci, cf, ci, cu, vp, vp]
lib.simulst_mma_train_bwd.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, ci, vp, ci, ci, ci, ci, cf, ci, cu, vp]
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
N, T = 512, 128


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]


def case(name, S, lens, flags=3 | 16):
    g = torch.Generator().manual_seed(1)
    dt = torch.bfloat16
    p = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, dt)
    e = torch.randn(N, T, S, generator=g).to(dev, dt)
    alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
    side = torch.zeros(N, T, 2, device=dev)
    ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
    gp = torch.empty_like(p); ge = torch.empty_like(e)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    mask = None
    if lens is not None:
        mask = (torch.arange(S)[None, :] >= lens[:, None]).to(dev).view(torch.uint8).contiguous()
    else:
        flags &= ~16
    mp = mask.data_ptr() if mask is not None else None

    def fwd():
        rc = lib.simulst_mma_train_fwd(p.data_ptr(), 1, e.data_ptr(), 1, mp, alpha.data_ptr(), beta.data_ptr(),
                                       side.data_ptr(), N, T, S, 1e-6, 0, flags, status.data_ptr(), st)
        assert rc == 0, rc

    def bwd():
        rc = lib.simulst_mma_train_bwd(p.data_ptr(), 1, e.data_ptr(), 1, mp, alpha.data_ptr(), side.data_ptr(),
                                       ga.data_ptr(), gb.data_ptr(), gp.data_ptr(), 1, ge.data_ptr(), 1,
                                       N, T, S, 1e-6, 0, flags, st)
        assert rc == 0, rc

    fwd(); bwd(); torch.cuda.synchronize()
    f, b = timeit(fwd), timeit(bwd)
    print(json.dumps({"lib": os.path.basename(os.path.dirname(os.path.abspath(path))) + "/" + os.path.basename(path),
                      "case": name, "S": S, "fwd_us": round(f, 1), "bwd_us": round(b, 1), "sum_us": round(f + b, 1),
                      "status": int(status.item())}), flush=True)


S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator().manual_seed(1236)
case("nomask", S, None)
case("full", S, torch.full((N,), S, dtype=torch.long))
case("uniform_half_to_full", S, torch.randint(S // 2, S + 1, (N,), generator=g))
case("fixed_3q", S, torch.full((N,), 3 * S // 4, dtype=torch.long))
case("fixed_3q_plus3", S, torch.full((N,), 3 * S // 4 + 3, dtype=torch.long))
