import torch, time
n=64*65536
d=torch.randn(n,dtype=torch.float64,device='cuda'); h=torch.empty(n,dtype=torch.float64).pin_memory()
s=torch.cuda.Stream()
for name,src in [('1d',d)]:
    with torch.cuda.stream(s):
        for _ in range(3): h.copy_(src,non_blocking=True)
        s.synchronize(); t0=time.perf_counter()
        for _ in range(20): h.copy_(src,non_blocking=True)
        s.synchronize(); t1=time.perf_counter()
    print(name, n*8*20/(t1-t0)/1e9,'GB/s')
# 2D strided
d2=torch.randn(64,65776,dtype=torch.float64,device='cuda'); h2=torch.empty(64,65536,dtype=torch.float64).pin_memory()
with torch.cuda.stream(s):
    for _ in range(3): h2.copy_(d2[:,240:240+65536],non_blocking=True)
    s.synchronize(); t0=time.perf_counter()
    for _ in range(20): h2.copy_(d2[:,240:240+65536],non_blocking=True)
    s.synchronize(); t1=time.perf_counter()
print('2d', n*8*20/(t1-t0)/1e9,'GB/s')
# H2D
with torch.cuda.stream(s):
    for _ in range(3): d.copy_(h,non_blocking=True)
    s.synchronize(); t0=time.perf_counter()
    for _ in range(20): d.copy_(h,non_blocking=True)
    s.synchronize(); t1=time.perf_counter()
print('h2d', n*8*20/(t1-t0)/1e9,'GB/s')
