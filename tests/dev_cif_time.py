"""Development timing probe (not a test): per-call device time of the CIF C-ABI entry points at
the C3 shape (B=64, S=1500, C=256), L2 flushed between repetitions."""
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib
from simulst_b200.models.torch_cif import cif_function

lib = _lib.load()
dev = torch.device("cuda")
B, S, C = 64, 1500, 256
dt = torch.float32
if len(sys.argv) > 3:
    B, S, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4:
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[4]]
g = torch.Generator().manual_seed(2024)
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
gx = torch.empty_like(x); ga = torch.empty_like(a); ws = torch.empty(2 * B * S, device=dev)
seg = torch.empty(B, T + 2, dtype=torch.int32, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
XD = _lib.dtype_enum(dt)

def plan():
    return lib.simulst_cif_plan(a.data_ptr(), 0, None, desired.data_ptr(), tl.data_ptr(), csum.data_ptr(),
                                scale.data_ptr(), asum.data_ptr(), len0.data_ptr(), cnt.data_ptr(), seg.data_ptr(), T + 2, B, S, 1.0,
                                status.data_ptr(), st)
def fwd():
    return lib.simulst_cif_fwd(x.data_ptr(), XD, csum.data_ptr(), scale.data_ptr(), a.data_ptr(), 0, None,
                               seg.data_ptr(), T + 2, out.data_ptr(), delays.data_ptr(), None, len0.data_ptr(), None, None,
                               B, S, C, T, T, 1.0, 0.5, 1, st)
def bwd():
    return lib.simulst_cif_bwd(x.data_ptr(), XD, csum.data_ptr(), scale.data_ptr(), a.data_ptr(), 0, None,
                               go.data_ptr(), gd.data_ptr(), None, None, None, asum.data_ptr(), None,
                               gx.data_ptr(), ga.data_ptr(), ws.data_ptr(), B, S, C, T, T, 1.0, 0.5, 1, st)

def timeit(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(); e1.record(); torch.cuda.synchronize()
        assert rc == 0, rc
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]

import os
if os.environ.get("TILE"):
    fc, fr = [int(v) for v in os.environ["TILE"].split(",")]
    assert lib.simulst_cif_set_tile_rows(fc, fr) == 0
es = x.element_size()
plan(); fwd(); bwd(); torch.cuda.synchronize()
# back-to-back chain (what bench.py times): plan + fwd + bwd, no flush in between
def chain():
    plan(); fwd(); return bwd()
tc = timeit(chain)
print(f"TILE={os.environ.get('TILE','auto')} chain {tc:.1f} us", flush=True)
tp, tf, tb = timeit(plan), timeit(fwd), timeit(bwd)
bf = B * S * C * es + B * T * C * es + B * S * 8 + B * T * es
bb = 2 * B * S * C * es + B * T * C * es + B * S * 8
print(f"B={B} S={S} C={C} T={T} {dt}: plan {tp:.1f} us | fwd {tf:.1f} us ({bf/tf/1e3:.0f} GB/s) | "
      f"bwd {tb:.1f} us ({bb/tb/1e3:.0f} GB/s) | sum {tp+tf+tb:.1f} us "
      f"frac={(bf+bb)/(tp+tf+tb)/1e3/6549.8:.3f}", flush=True)

# through the Python API (includes allocations and the reference's host read of T)
xr = x.clone().requires_grad_(); ar = a.clone().requires_grad_()
def api():
    xr.grad = None; ar.grad = None
    r = cif_function(xr, ar, beta=1.0, tail_thres=0.5, target_lengths=tl)
    torch.autograd.backward([r["cif_out"][0], r["delays"][0]], [go, gd])
for _ in range(3): api()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): api()
e1.record(); torch.cuda.synchronize()
print(f"python API fwd+bwd: {e0.elapsed_time(e1)/20*1e3:.1f} us")
