import torch, time
n=100*1024*1024//8
h=torch.empty(n,dtype=torch.float64).pin_memory(); d=torch.empty(n,dtype=torch.float64,device='cuda')
for name,fn in (("H2D",lambda: d.copy_(h,non_blocking=True)),("D2H",lambda: h.copy_(d,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, "pinned %.1f GB/s"%(n*8/dt/1e9))
p=torch.empty(n,dtype=torch.float64)
t=time.perf_counter(); d.copy_(p); torch.cuda.synchronize(); print("H2D pageable %.1f GB/s"%(n*8/(time.perf_counter()-t)/1e9))
