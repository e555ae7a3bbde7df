import torch, time
dev=torch.device('cuda',0)
for mb in (13, 64, 256):
    n=mb*1024*1024
    d=torch.empty(n,dtype=torch.uint8,device=dev)
    h=torch.empty(n,dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(20): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('D2H DMA %d MB: %.1f GB/s'%(mb, 20*n/dt/1e9))
    t0=time.perf_counter()
    for _ in range(20): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('H2D DMA %d MB: %.1f GB/s'%(mb, 20*n/dt/1e9))
import subprocess
print(subprocess.run(['nvidia-smi','--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current','--format=csv'],capture_output=True,text=True).stdout)
