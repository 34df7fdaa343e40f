"""Development probe: where does the Python-side time of cif_function + backward go?"""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from simulst_b200.models.torch_cif import cif_function
dev = "cuda"
g = torch.Generator().manual_seed(2024)
b, s, c = 64, 1500, 256
x = torch.randn(b, s, c, generator=g).to(dev).requires_grad_()
a = torch.sigmoid(torch.randn(b, s, generator=g) - 1.0).to(dev).requires_grad_()
tl = a.detach().sum(1).round().clamp(min=1).long().cpu()
res = cif_function(x, a, beta=1.0, tail_thres=0.5, target_lengths=tl)
go = torch.randn_like(res["cif_out"][0]); gd = torch.randn_like(res["delays"][0])
def step():
    x.grad = None; a.grad = None
    r = cif_function(x, a, beta=1.0, tail_thres=0.5, target_lengths=tl)
    torch.autograd.backward([r["cif_out"][0], r["delays"][0]], [go, gd])
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize()
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28); print(st.getvalue()[:6000])
