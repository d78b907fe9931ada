"""GPU probe: FP64 peaks (cuBLAS DGEMM, copy bandwidth) and eigensolver timings. Writes gpurun_out/probe.json."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200._engine import Eigh, to_dev, _p
from gglasso_b200 import _lib
from gglasso_b200.datagen import synthetic_mgl

out = {}
dev = torch.device("cuda")
def tm(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

# FP64 DGEMM peak (cuBLAS) -- roofline denominator for the DMMA kernels
for n in (2048, 4096, 8192):
    A=torch.randn(n,n,dtype=torch.float64,device=dev); B=torch.randn(n,n,dtype=torch.float64,device=dev)
    best,med=tm(lambda: torch.matmul(A,B))
    out[f"dgemm_{n}_tflops_best"]=2*n**3/best/1e9; out[f"dgemm_{n}_tflops_med"]=2*n**3/med/1e9
A=torch.randn(20,1000,1000,dtype=torch.float64,device=dev)
best,med=tm(lambda: torch.matmul(A,A.transpose(1,2)))
out["dgemm_batched_20x1000_tflops"]=2*20*1e9/best/1e9
x=torch.empty(1<<28,dtype=torch.float64,device=dev); y=torch.empty_like(x)
best,med=tm(lambda: y.copy_(x))
out["copy_gbs"]=2*x.numel()*8/best/1e6
# torch (cuSOLVER) eigh for context
S=synthetic_mgl(20,1000,N=2000,seed=1,kind="fused")
W=np.eye(1000)[None]-S
Wd=to_dev(W,dev)
best,med=tm(lambda: torch.linalg.eigh(Wd), n=2, warm=1)
out["torch_eigh_20x1000_ms"]=best
lib=_lib.load()
for nb2 in (32,64,128):
    e=Eigh(20,1000,dev); e.nb2=nb2
    def run():
        A=Wd.clone(); e.eigh(A)
    best,med=tm(run,n=3,warm=1)
    out[f"gg_eigh_20x1000_nb2_{nb2}_ms"]=best; out[f"gg_eigh_20x1000_nb2_{nb2}_sweeps"]=e.sweeps[-1]
    A=Wd.clone(); D=e.eigh(A).clone()
    Dref=torch.linalg.eigvalsh(Wd)
    out[f"gg_eigh_nb2_{nb2}_eigerr"]=float((torch.sort(D,1)[0]-Dref).abs().max())
    O=torch.empty_like(A)
    best,med=tm(lambda: e.recon(A,O,0),n=5,warm=1)
    out[f"gg_recon_20x1000_ms"]=best
for (M,p) in ((10,500),(1,100),(5,100),(1,1289),(1,2000)):
    Wd2=to_dev(np.eye(p)[None]-synthetic_mgl(M,p,N=2*p,seed=2),dev)
    e=Eigh(M,p,dev)
    def run2():
        A=Wd2.clone(); e.eigh(A)
    best,med=tm(run2,n=3,warm=1)
    out[f"gg_eigh_{M}x{p}_ms"]=best; out[f"gg_eigh_{M}x{p}_sweeps"]=e.sweeps[-1]
    best,med=tm(lambda: torch.linalg.eigh(Wd2), n=2, warm=1)
    out[f"torch_eigh_{M}x{p}_ms"]=best
os.makedirs("gpurun_out",exist_ok=True)
json.dump(out,open("gpurun_out/probe.json","w"),indent=1)
print(json.dumps(out,indent=1))
