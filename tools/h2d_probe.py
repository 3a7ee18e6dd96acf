import torch, time
for mb in (0.5, 4.7, 64):
    n=int(mb*1e6); h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device="cuda")
    for _ in range(3): d.copy_(h,non_blocking=True)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): d.copy_(h,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t=e0.elapsed_time(e1)/20*1e-3
    print(mb,"MB", round(t*1e6,1),"us", round(n/t/1e9,1),"GB/s")
