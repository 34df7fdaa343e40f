"""Long-form sweep (BASELINE config 5, SURVEY 8d "C5"): MMA fused forward+backward over
src in {512..6000} x tgt in {64..512} and CIF forward+backward over the same source lengths,
through the C ABI on resident buffers.  Prints one JSON line per point with achieved algorithmic
GB/s and the fraction of the measured HBM peak.  Not a test; run on a B200:
    python tests/dev_sweep.py [rows] > profiles/<round>_sweep.jsonl"""
import json
import os
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]


def mma_point(N, T, S):
    g = torch.Generator().manual_seed(1234)
    dt = torch.bfloat16
    p = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, dt)
    e = torch.randn(N, T, S, generator=g).to(dev, dt)
    # inputs dense (what a caller hands over); outputs with the pitch the library asks for (= S whenever
    # dense rows are already 16-byte multiples), as the Python wrapper allocates them
    ld = S if os.environ.get("DENSE_OUT") else int(lib.simulst_mma_out_pitch(S))
    alpha = torch.empty(N, T, ld, device=dev); beta = torch.empty_like(alpha)
    side = torch.empty(N, T, 2, device=dev)
    ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
    gp = torch.empty(N, T, ld, device=dev, dtype=dt); ge = torch.empty_like(gp)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def fwd():
        rc = lib.simulst_mma_train_fwd_pitched(p.data_ptr(), 1, S, e.data_ptr(), 1, S, None, alpha.data_ptr(), ld,
                                               beta.data_ptr(), ld, side.data_ptr(), None, N, T, S, 1e-6, 0, 3,
                                               status.data_ptr(), st)
        assert rc == 0, rc

    def bwd():
        rc = lib.simulst_mma_train_bwd_pitched(p.data_ptr(), 1, S, e.data_ptr(), 1, S, None, alpha.data_ptr(), ld,
                                               side.data_ptr(), ga.data_ptr(), S, gb.data_ptr(), S, None,
                                               gp.data_ptr(), 1, ld, ge.data_ptr(), 1, ld, N, T, S, 1e-6, 0, 3, st)
        assert rc == 0, rc

    fwd(); bwd(); torch.cuda.synchronize()
    f, b = timeit(fwd), timeit(bwd)
    el = N * T * S
    print(json.dumps({"op": "mma_fwd_bwd", "rows": N, "tgt": T, "src": S, "dtype_in": "bf16",
                      "out_pitch": ld, "fwd_us": round(f, 1), "bwd_us": round(b, 1),
                      "elements_per_s": el / (f + b) * 1e6,
                      "fwd_gbs": round(el * 12 / f / 1e3, 1), "bwd_gbs": round(el * 20 / b / 1e3, 1),
                      "fwd_bwd_gbs": round(el * 32 / (f + b) / 1e3, 1),
                      "frac_of_hbm_peak": round(el * 32 / (f + b) / 1e3 / PEAK, 3)}), flush=True)


def cif_point(B, S, C=256):
    g = torch.Generator().manual_seed(2024)
    dt = torch.float32
    x = torch.randn(B, S, C, generator=g).to(dev, dt)
    a = torch.sigmoid(torch.randn(B, S, generator=g) - 1.0).to(dev)
    tl = a.sum(1).round().clamp(min=1).long()
    T = int(tl.max())
    desired = (1.0 * tl.to(dt) + 1e-4).float()
    csum = torch.empty(B, S, device=dev); scale = torch.empty(B, device=dev); asum = torch.empty(B, device=dev)
    len0 = torch.empty(B, dtype=torch.int64, device=dev)
    cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    out = torch.empty(B, T, C, device=dev, dtype=dt); delays = torch.empty(B, T, device=dev, dtype=dt)
    go = torch.randn(B, T, C, device=dev, dtype=dt); gd = torch.randn(B, T, device=dev, dtype=dt)
    gx = torch.empty_like(x); gal = torch.empty_like(a); ws = torch.empty(2 * B * S, device=dev)
    seg = torch.empty(B, T + 2, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def step():
        rc = lib.simulst_cif_plan(a.data_ptr(), 0, None, desired.data_ptr(), tl.data_ptr(), csum.data_ptr(),
                                  scale.data_ptr(), asum.data_ptr(), len0.data_ptr(), cnt.data_ptr(), seg.data_ptr(),
                                  T + 2, B, S, 1.0, status.data_ptr(), st)
        assert rc == 0, rc
        rc = lib.simulst_cif_fwd(x.data_ptr(), 0, csum.data_ptr(), scale.data_ptr(), a.data_ptr(), 0, None,
                                 seg.data_ptr(), T + 2, out.data_ptr(), delays.data_ptr(), None, len0.data_ptr(),
                                 None, None, B, S, C, T, T, 1.0, 0.5, 1, st)
        assert rc == 0, rc
        rc = lib.simulst_cif_bwd(x.data_ptr(), 0, csum.data_ptr(), scale.data_ptr(), a.data_ptr(), 0, None,
                                 go.data_ptr(), gd.data_ptr(), None, None, None, asum.data_ptr(), None,
                                 gx.data_ptr(), gal.data_ptr(), ws.data_ptr(), B, S, C, T, T, 1.0, 0.5, 1, st)
        assert rc == 0, rc

    step(); torch.cuda.synchronize()
    t = timeit(step)
    alg = B * S * (C * 4 * (3 + 2 * T / S) + 8)
    print(json.dumps({"op": "cif_fwd_bwd", "B": B, "src": S, "C": C, "T": T, "dtype": "f32",
                      "us": round(t, 1), "frames_per_s": B * S / t * 1e6,
                      "gbs": round(alg / t / 1e3, 1), "frac_of_hbm_peak": round(alg / t / 1e3 / PEAK, 3)}), flush=True)


if __name__ == "__main__":
    SRC = [int(v) for v in os.environ["SRC"].split(",")] if os.environ.get("SRC") else (512, 1024, 2048, 4096, 6000)
    for S in SRC:
        for T in ([int(v) for v in os.environ["TGT"].split(",")] if os.environ.get("TGT") else (64, 128, 256, 512)):
            mma_point(ROWS, T, S)
            torch.cuda.empty_cache()
    for S in (() if os.environ.get("NO_CIF") else (512, 1024, 1500, 2048, 4096, 6000)):
        cif_point(64, S)
        torch.cuda.empty_cache()
