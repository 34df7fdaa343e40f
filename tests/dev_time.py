"""Development timing probe (not a test): C2-shaped fwd / bwd timings per kernel config."""
import os
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
N, T, S = 512, 128, 1024
if len(sys.argv) > 3:
    N, T, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(1234)
dt = torch.bfloat16
p = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2).to(dev, dt)
e = torch.randn(N, T, S, generator=g).to(dev, dt)
alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
side = torch.empty(N, T, 2, device=dev)
ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
gp = torch.empty_like(p); ge = torch.empty_like(e)
status = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
mask = None
if os.environ.get('MASK'):
    lens = torch.randint(S // 2, S + 1, (N,), generator=g)
    mask = (torch.arange(S)[None, :] >= lens[:, None]).to(dev).view(torch.uint8).contiguous()
mp_ = mask.data_ptr() if mask is not None else None
flags = int(os.environ.get('FLAGS', '3'))

def fwd():
    return lib.simulst_mma_train_fwd(p.data_ptr(), 1, e.data_ptr(), 1, mp_, alpha.data_ptr(), beta.data_ptr(),
                                     side.data_ptr(), N, T, S, 1e-6, 0, flags, status.data_ptr(), st)
def bwd():
    return lib.simulst_mma_train_bwd(p.data_ptr(), 1, e.data_ptr(), 1, mp_, alpha.data_ptr(), side.data_ptr(),
                                     ga.data_ptr(), gb.data_ptr() if flags & 1 else None, gp.data_ptr(), 1, ge.data_ptr(), 1,
                                     N, T, S, 1e-6, 0, flags, st)

def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); rc = fn(); b.record(); torch.cuda.synchronize()
        assert rc == 0, rc
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]

el = N * T * S
import os
cfgs = [tuple(map(int, c.split('x'))) for c in os.environ.get('CFGS', '128x8').split(',')]
for cfg in cfgs:
    if cfg[0] * cfg[1] < S:
        continue
    modes = [int(m) for m in os.environ["MODES"].split(",")] if os.environ.get("MODES") else ([0, 1, 2, 3] if os.environ.get("BOTH") else [1])
    ref = None
    for tma in modes:
        assert lib.simulst_mma_set_config(*cfg) == 0
        lib.simulst_mma_set_pipeline(tma)
        fwd(); gp.zero_(); ge.zero_(); bwd(); torch.cuda.synchronize()
        if ref is None:
            ref = (gp.float().clone(), ge.float().clone())
        else:
            dp = (gp.float() - ref[0]).abs().max().item(); de = (ge.float() - ref[1]).abs().max().item()
            print(f"   vs first mode: max|d grad_p| {dp:.3e} (scale {ref[0].abs().max().item():.3e})  max|d grad_e| {de:.3e} (scale {ref[1].abs().max().item():.3e})"
                  f"  equal: {torch.equal(gp.float(), ref[0])} {torch.equal(ge.float(), ref[1])}")
        f_min, f_med = timeit(fwd)
        b_min, b_med = timeit(bwd)
        print(f"cfg={cfg} pipe={tma}: fwd {f_med:8.1f} us ({el*12/f_med/1e3:7.1f} GB/s)  "
              f"bwd {b_med:8.1f} us ({el*20/b_med/1e3:7.1f} GB/s)  "
              f"fwd+bwd {el/(f_med+b_med)*1e6/1e9:7.2f} Gelem/s  frac={el*32/(f_med+b_med)/1e3/6540.2:.3f}",
              flush=True)
lib.simulst_mma_set_config(0, 0); lib.simulst_mma_set_pipeline(1)
