import torch, time
x = torch.empty(64*1024*1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D pinned 64MiB GB/s', x.numel()/dt/1e9)
y = torch.empty(64*1024*1024, dtype=torch.uint8)
torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(y); torch.cuda.synchronize(); print('H2D pageable GB/s', y.numel()/(time.perf_counter()-t)/1e9)
