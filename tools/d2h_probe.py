import torch, time
n=4_800_000_000
d=torch.empty(n,dtype=torch.uint8,device='cuda'); h=torch.empty(n,dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(d,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(3): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
print("1 stream 1D: %.1f ms %.1f GB/s"%(dt*1e3,n/dt/1e9))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
half=n//2
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): h[:half].copy_(d[:half],non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(d[half:],non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
print("2 streams: %.1f ms %.1f GB/s"%(dt*1e3,n/dt/1e9))
# h2d concurrently with d2h
d2=torch.empty(82_000_000,dtype=torch.uint8,device='cuda'); h2=torch.empty(82_000_000,dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
print("d2h + concurrent h2d: %.1f ms"%(dt*1e3))
