import torch, time
x = torch.empty(640*1024*1024//2, dtype=torch.bfloat16).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1,2,4):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print('h2d full', x.numel()*2/dt/1e9, 'GB/s')
# chunked on two streams
s=[torch.cuda.Stream(),torch.cuda.Stream()]
ch=49*1024*1024//2
torch.cuda.synchronize(); t=time.perf_counter()
for i in range(0,x.numel(),ch):
    with torch.cuda.stream(s[(i//ch)&1]): d[i:i+ch].copy_(x[i:i+ch], non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('h2d chunked 2 streams', x.numel()*2/dt/1e9)
y=torch.empty(22*1024*1024, dtype=torch.uint8).pin_memory(); dd=torch.empty_like(y, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter(); y.copy_(dd, non_blocking=True); torch.cuda.synchronize(); print('d2h', y.numel()/(time.perf_counter()-t)/1e9)
import subprocess; print(subprocess.run(['nvidia-smi','--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max','--format=csv'],capture_output=True,text=True).stdout)
