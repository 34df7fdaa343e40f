"""Development probe (not a test): host-buffer pipeline (e2e leg of bench.py) over chunk / stream counts."""
import sys
import torch
sys.path.insert(0, ".")
from simulst_b200.host_pipeline import MMAHostPipeline

N, T, S = 512, 128, 1024
dev = torch.device("cuda")
dt = torch.bfloat16
g = torch.Generator().manual_seed(1)
p_host = torch.sigmoid(torch.randn(N, T, S, generator=g) - 2.0).to(dt).pin_memory()
e_host = torch.randn(N, T, S, generator=g).to(dt).pin_memory()
gp_host = torch.empty(N, T, S, dtype=dt).pin_memory()
ge_host = torch.empty(N, T, S, dtype=dt).pin_memory()
ga = torch.randn(N, T, S, device=dev) * 1e-2
gb = torch.randn(N, T, S, device=dev)
for chunks in (8, 16, 32, 64):
    for streams in (1, 2, 4, 8):
        pipe = MMAHostPipeline(N, T, S, dtype=dt, device=dev, chunks=chunks, compute_streams=streams)
        for _ in range(2):
            pipe.step(p_host, e_host, ga, gb, gp_host, ge_host)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            pipe.step(p_host, e_host, ga, gb, gp_host, ge_host)
        b.record(); torch.cuda.synchronize()
        print(f"chunks {chunks:3d} streams {streams}: {a.elapsed_time(b) / 5:.3f} ms/step", flush=True)
        del pipe
