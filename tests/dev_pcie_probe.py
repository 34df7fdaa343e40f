import torch, time
dev=torch.device('cuda')
n=268435456
h_in=torch.empty(n,dtype=torch.uint8).pin_memory(); h_out=torch.empty(n,dtype=torch.uint8).pin_memory()
d_in=torch.empty(n,dtype=torch.uint8,device=dev); d_out=torch.empty(n,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn,reps=5):
    fn(); torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        t0=time.perf_counter(); fn(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
def both(): h2d(); d2h()
a=t(h2d); b=t(d2h); c=t(both)
print(f"H2D {n/a/1e9:.1f} GB/s ({a*1e3:.2f} ms)  D2H {n/b/1e9:.1f} GB/s ({b*1e3:.2f} ms)  both concurrently {c*1e3:.2f} ms ({2*n/c/1e9:.1f} GB/s aggregate)")
