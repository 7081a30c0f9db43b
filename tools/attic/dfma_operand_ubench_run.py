import numpy as np, torch
from cuda.bindings import driver as cu
def ck(r):
    assert r[0]==cu.CUresult.CUDA_SUCCESS, r[0]; return r[1] if len(r)==2 else r[1:]
torch.zeros(1,device='cuda')
data=np.fromfile('build/ubench/ub.cubin',dtype=np.uint8); mod=ck(cu.cuModuleLoadData(data.ctypes.data))
out=torch.zeros(148*8*256,dtype=torch.float64,device='cuda')
for name in ['leaf2','leaf4','leaf8','leaf16','grouped2','grouped4','gonly','vonly2','vonly4','vonly16']:
    fn=ck(cu.cuModuleGetFunction(mod,('ub_'+name).encode()))
    iters=4000
    res=[]
    for bps in (1,2,4):
        args=(np.array([out.data_ptr()],dtype=np.uint64),np.array([iters],dtype=np.uint32))
        argp=np.array([a.ctypes.data for a in args],dtype=np.uint64)
        st=torch.cuda.current_stream().cuda_stream
        def launch(): ck(cu.cuLaunchKernel(fn,148*bps,1,1,256,1,1,0,st,argp.ctypes.data,0))
        launch(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)
        nfma=148*bps*256*iters*64
        res.append('w/smsp %d: %.3f'%(bps*2, nfma/32/(ms*1e-3)/(148*2*1.965e9)))
    print(name.ljust(10),'  '.join(res))
