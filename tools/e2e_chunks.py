import sys, time, math, torch
sys.path.insert(0, '/root/repo')
import equilib_b200
from equilib_b200.pipeline import HostPipeline
dev=torch.device('cuda',0)
b=32
src=torch.rand((b,3,2048,4096),device=dev).to(torch.bfloat16)
rots=[{"roll":0.01*i,"pitch":0.4*math.sin(i),"yaw":0.05*i} for i in range(b)]
op=equilib_b200.Equi2Equi()
hin=torch.empty(src.shape,dtype=src.dtype,pin_memory=True); hin.copy_(src)
hout=torch.empty(src.shape,dtype=src.dtype,pin_memory=True)
for chunk in (1,2,4,8):
    pipe=HostPipeline(op,chunk=chunk,device=dev)
    pipe(hin,rots,out=hout); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(5): pipe(hin,rots,out=hout)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/5
    print(chunk, dt*1e3, "ms", b*2048*4096/dt/1e6, "Mpix/s", flush=True)
