import torch, time
dev=torch.device('cuda:0')
n=2<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n//2,dtype=torch.uint8).pin_memory()
d2=torch.empty(n//2,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=3):
    best=1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0=time.perf_counter(); f(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
a=t(lambda: h.copy_(d,non_blocking=True)); print("D2H 2GiB", n/a/1e9,"GB/s")
b=t(lambda: d.copy_(h,non_blocking=True)); print("H2D 2GiB", n/b/1e9,"GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
c=t(both); print("D2H 2GiB + H2D 1GiB concurrently", c*1e3,"ms -> D2H", n/c/1e9)
for sz in (64<<20, 256<<20):
    hh=h[:sz]; dd=d[:sz]
    a=t(lambda: hh.copy_(dd,non_blocking=True)); print("D2H",sz>>20,"MiB", sz/a/1e9)
