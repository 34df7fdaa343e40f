"""Development probe: one fwd + one bwd launch at the C2 shape for ncu capture."""
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
N, T, S = 512, 128, 1024
dt = torch.bfloat16
p = torch.sigmoid(torch.randn(N, T, S, device=dev) - 2).to(dt)
e = torch.randn(N, T, S, device=dev).to(dt)
alpha = torch.empty(N, T, S, device=dev); beta = torch.empty_like(alpha)
side = torch.empty(N, T, 2, device=dev)
ga = torch.randn(N, T, S, device=dev) * 0.01; gb = torch.randn(N, T, S, device=dev)
gp = torch.empty_like(p); ge = torch.empty_like(e)
status = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
import os
lib.simulst_mma_set_pipeline(int(os.environ.get("PIPE", "1")))
for _ in range(2):
    lib.simulst_mma_train_fwd(p.data_ptr(), 1, e.data_ptr(), 1, None, alpha.data_ptr(), beta.data_ptr(),
                              side.data_ptr(), N, T, S, 1e-6, 0, 3, status.data_ptr(), st)
    lib.simulst_mma_train_bwd(p.data_ptr(), 1, e.data_ptr(), 1, None, alpha.data_ptr(), side.data_ptr(),
                              ga.data_ptr(), gb.data_ptr(), gp.data_ptr(), 1, ge.data_ptr(), 1,
                              N, T, S, 1e-6, 0, 3, st)
torch.cuda.synchronize()
