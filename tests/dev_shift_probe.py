"""Where does the time of the SHIFT (unaligned-row) dense kernels go?  Times fwd / bwd at 512 rows x 128
steps for: S=1504 dense; S=1504 with an all-live right-padding mask + promise (MASKED instantiation);
S=1504 from a base pointer that is 8 bytes off (SHIFT, no partial thread); S=1500 (SHIFT + partial thread).
Not a test.  python tests/dev_shift_probe.py [case]"""
import json
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
N, T = 512, 128
ONLY = sys.argv[1] if len(sys.argv) > 1 else None
REPS = 1 if ONLY else 5


def timeit(fn, reps=REPS):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]


def case(name, S, off_elems=0, masked=False, lens=None, base_flags=3):
    if ONLY and ONLY != name:
        return
    g = torch.Generator().manual_seed(1)
    dt = torch.bfloat16
    ld = int(lib.simulst_mma_out_pitch(S))

    def at_offset(x):
        buf = torch.empty(x.numel() + 64, dtype=x.dtype, device=dev)
        v = buf[off_elems:off_elems + x.numel()].view(x.shape)
        v.copy_(x)
        return v
    p = at_offset(torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, dt))
    e = at_offset(torch.randn(N, T, S, generator=g).to(dev, dt))
    alpha = torch.empty(N, T, ld, device=dev); beta = torch.empty_like(alpha)
    side = torch.zeros(N, T, 2, device=dev)
    ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
    gp = torch.empty(N, T, ld, device=dev, dtype=dt); ge = torch.empty_like(gp)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    mask = None
    flags = base_flags
    if masked:
        L = torch.full((N,), S, dtype=torch.long) if lens is None else lens
        mask = (torch.arange(S)[None, :] >= L[:, None]).to(dev).view(torch.uint8).contiguous()
        flags |= 16
    mp = mask.data_ptr() if mask is not None else None

    def fwd():
        rc = lib.simulst_mma_train_fwd_pitched(p.data_ptr(), 1, S, e.data_ptr(), 1, S, mp, alpha.data_ptr(), ld,
                                               beta.data_ptr(), ld, side.data_ptr(), None, N, T, S, 1e-6, 0, flags,
                                               status.data_ptr(), st)
        assert rc == 0, rc

    soft = bool(flags & 2)

    def bwd():
        rc = lib.simulst_mma_train_bwd_pitched(p.data_ptr(), 1, S, e.data_ptr(), 1, S, mp, alpha.data_ptr(), ld,
                                               side.data_ptr(), ga.data_ptr(), S, gb.data_ptr() if soft else None, S, None,
                                               gp.data_ptr(), 1, ld, ge.data_ptr() if soft else None, 1, ld, N, T, S,
                                               1e-6, 0, flags, st)
        assert rc == 0, rc

    fwd(); bwd(); torch.cuda.synchronize()
    f, b = timeit(fwd), timeit(bwd)
    print(json.dumps({"case": name, "S": S, "fwd_us": round(f, 1), "bwd_us": round(b, 1), "status": int(status.item())}),
          flush=True)


case("dense1504", 1504)
case("masked1504", 1504, masked=True)
case("shift1504_off8B", 1504, off_elems=4)
case("shift1500", 1500)
case("dense1024", 1024)
case("masked1024_full", 1024, masked=True)
case("shift1024_off8B", 1024, off_elems=4)
case("shift1000_off2B", 1000, off_elems=1)
case("dense1504_nomp", 1504, base_flags=2)
case("masked1504_nomp", 1504, masked=True, base_flags=2)
case("dense1504_hard", 1504, base_flags=1)
case("masked1504_hard", 1504, masked=True, base_flags=1)
