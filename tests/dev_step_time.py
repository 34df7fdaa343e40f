import sys, json, torch
sys.path.insert(0, ".")
import bench
from simulst_b200 import _lib
lib=_lib.load(); dev=torch.device("cuda")
g = torch.Generator().manual_seed(3000)
r, s = 1024, 1024
p = torch.sigmoid(torch.randn(r, s, generator=g) - 2.0).to(dev); se = torch.randn(r, s, generator=g).to(dev)
print(json.dumps(bench.step_variants(lib, dev, p, se)))
