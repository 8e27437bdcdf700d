import torch, time
for mb in (9.6, 38.4, 256):
    n=int(mb*1e6)
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
    for _ in range(3): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): d.copy_(h,non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b)/20
    print("H2D %.1f MB: %.3f ms  %.1f GB/s"%(mb,ms,n/ms/1e6))
    a.record()
    for _ in range(20): h.copy_(d,non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b)/20
    print("D2H %.1f MB: %.3f ms  %.1f GB/s"%(mb,ms,n/ms/1e6))
