"""Development timing probe (not a test): dense vs pooled p_choose at the training shape
(fixed pre-decision ratio 8, exp/2-mma.sh:56-57), fwd and bwd, through the C ABI."""
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
N, T, S, R = 512, 128, 1024, 8
if len(sys.argv) > 4:
    N, T, S, R = [int(v) for v in sys.argv[1:5]]
SP = (S + R - 1) // R
g = torch.Generator().manual_seed(1234)
dt = torch.bfloat16
pp = torch.sigmoid(torch.randn(N, T, SP, generator=g) - 1).to(dev, dt)
e = torch.randn(N, T, S, generator=g).to(dev, dt)
pd = torch.zeros(N, T, S, device=dev, dtype=dt)
alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
alpha2 = torch.empty_like(alpha); beta2 = torch.empty_like(alpha)
side = torch.empty(N, T, 2, device=dev)
ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
gpp = torch.empty_like(pp); gp = torch.empty_like(pd); ge = torch.empty_like(e); ge2 = torch.empty_like(e)
status = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
flags = 3
ws = torch.empty(int(lib.simulst_mma_pooled_workspace_bytes(N, T, S, R)), dtype=torch.uint8, device=dev)

def fwd_pooled(dense):
    return lib.simulst_mma_train_fwd_pooled(pp.data_ptr(), 1, R, e.data_ptr(), 1, None, pd.data_ptr() if dense else None,
                                            alpha.data_ptr(), beta.data_ptr(), side.data_ptr(), None, ws.data_ptr(), N, T, S,
                                            1e-6, 0, flags, status.data_ptr(), st)
def bwd_pooled():
    return lib.simulst_mma_train_bwd_pooled(pp.data_ptr(), 1, R, e.data_ptr(), 1, None, None, alpha.data_ptr(),
                                            side.data_ptr(), ga.data_ptr(), gb.data_ptr(), None, gpp.data_ptr(), 1, None,
                                            ge.data_ptr(), 1, ws.data_ptr(), N, T, S, 1e-6, 0, flags, st)
delays = torch.empty(N, T, device=dev); gd = torch.randn(N, T, device=dev) / S
def fwd_pooled_lean():
    return lib.simulst_mma_train_fwd_pooled(pp.data_ptr(), 1, R, e.data_ptr(), 1, None, None, None, beta.data_ptr(),
                                            side.data_ptr(), delays.data_ptr(), ws.data_ptr(), N, T, S, 1e-6, 0, flags,
                                            status.data_ptr(), st)
def bwd_pooled_lean():
    return lib.simulst_mma_train_bwd_pooled(pp.data_ptr(), 1, R, e.data_ptr(), 1, None, None, None,
                                            side.data_ptr(), None, gb.data_ptr(), gd.data_ptr(), gpp.data_ptr(), 1, None,
                                            ge.data_ptr(), 1, ws.data_ptr(), N, T, S, 1e-6, 0, flags, st)
def fwd_dense_delays():
    return lib.simulst_mma_train_fwd_delays(pd.data_ptr(), 1, e.data_ptr(), 1, None, alpha2.data_ptr(), beta2.data_ptr(),
                                            side.data_ptr(), delays.data_ptr(), N, T, S, 1e-6, 0, flags, status.data_ptr(), st)
def bwd_dense_delays():
    return lib.simulst_mma_train_bwd_delays(pd.data_ptr(), 1, e.data_ptr(), 1, None, alpha2.data_ptr(), side.data_ptr(),
                                            None, gb.data_ptr(), gd.data_ptr(), gp.data_ptr(), 1, ge2.data_ptr(), 1,
                                            N, T, S, 1e-6, 0, flags, st)
def fwd_dense():
    return lib.simulst_mma_train_fwd(pd.data_ptr(), 1, e.data_ptr(), 1, None, alpha2.data_ptr(), beta2.data_ptr(),
                                     side.data_ptr(), N, T, S, 1e-6, 0, flags, status.data_ptr(), st)
def bwd_dense():
    return lib.simulst_mma_train_bwd(pd.data_ptr(), 1, e.data_ptr(), 1, None, alpha2.data_ptr(), side.data_ptr(),
                                     ga.data_ptr(), gb.data_ptr(), gp.data_ptr(), 1, ge2.data_ptr(), 1, N, T, S, 1e-6, 0,
                                     flags, st)

def timeit(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); rc = fn(); b.record(); torch.cuda.synchronize()
        assert rc == 0, rc
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]

assert fwd_pooled(True) == 0 and bwd_pooled() == 0 and fwd_dense() == 0 and bwd_dense() == 0
torch.cuda.synchronize()
print("fused:", lib.simulst_mma_pooled_is_fused(1, S, R, 0, flags, 0), "max|d| alpha/beta/ge:",
      float((alpha - alpha2).abs().max()), float((beta - beta2).abs().max()),
      float((ge.float() - ge2.float()).abs().max()), "status", int(status.item()))
el = N * T * S
tf0, tf1, tfd = timeit(lambda: fwd_pooled(False)), timeit(lambda: fwd_pooled(True)), timeit(fwd_dense)
tb, tbd = timeit(bwd_pooled), timeit(bwd_dense)
print(f"N{N} T{T} S{S} r{R}: fwd pooled {tf0:.1f} us (+dense out {tf1:.1f}) vs dense {tfd:.1f} us | "
      f"bwd pooled {tb:.1f} vs dense {tbd:.1f} us | step pooled {tf0 + tb:.1f} vs dense {tfd + tbd:.1f} us")
tfl, tbl = timeit(fwd_pooled_lean), timeit(bwd_pooled_lean)
tfdd, tbdd = timeit(fwd_dense_delays), timeit(bwd_dense_delays)
print(f"   latency-loss mode (no dense alpha, grad through expected delays): pooled fwd {tfl:.1f} + bwd {tbl:.1f} = "
      f"{tfl + tbl:.1f} us vs dense fwd {tfdd:.1f} + bwd {tbdd:.1f} = {tfdd + tbdd:.1f} us")
