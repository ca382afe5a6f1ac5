import torch, time
n=64*1024*1024
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,src,dst in (('D2H',d,h),('H2D',h,d)):
    for sz in (1<<20, 4<<20, 64<<20):
        torch.cuda.synchronize(); t=time.perf_counter()
        reps=max(4, (256<<20)//sz)
        for _ in range(reps): dst[:sz].copy_(src[:sz], non_blocking=True)
        torch.cuda.synchronize(); el=time.perf_counter()-t
        print(name, sz>>20,'MiB', round(reps*sz/el/1e9,1),'GB/s')
