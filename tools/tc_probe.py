import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from openrec_b200 import native as N
from oracle import openrec_oracle as O
eng=N.engine()
def dev(a): return torch.as_tensor(np.ascontiguousarray(a)).to('cuda',torch.float32)
for (B,inn,out) in [(512,256,128),(1111,479,1024),(4096,512,256),(777,1024,64),(256,32,128),(256,64,128)]:
    rng=np.random.default_rng(B)
    x,w=rng.standard_normal((B,inn)),rng.standard_normal((inn,out))*0.3
    tx,tw=dev(x),dev(w)
    x64,w64=tx.cpu().numpy().astype(np.float64),tw.cpu().numpy().astype(np.float64)
    ref=x64@w64
    ty=torch.empty(B,out,device='cuda'); eng.mlp_fwd(tx,tw,None,0,ty)
    got=ty.cpu().numpy().astype(np.float64)
    err=np.abs(got-ref); 
    f32=(tx.cpu()@tw.cpu()).numpy().astype(np.float64)
    print(f"fwd B={B} K={inn} N={out}: max abs err {err.max():.3e} (fp32 torch-cpu err {np.abs(f32-ref).max():.3e}) max|ref| {np.abs(ref).max():.2f} rel-to-rms {err.max()/np.sqrt((ref**2).mean()):.2e}")
    # dw = x^T dy : K = B
    dy=dev(rng.standard_normal((B,out)))
    dy64=dy.cpu().numpy().astype(np.float64)
    tdw=torch.empty_like(tw); tdx=torch.empty(B,inn,device='cuda')
    yy=torch.empty(B,out,device='cuda'); eng.mlp_fwd(tx,tw,None,0,yy)
    eng.mlp_bwd(tx,yy,tw,0,dy.clone(),tdx,tdw,None)
    e1=np.abs(tdw.cpu().numpy()-x64.T@dy64).max(); e2=np.abs(tdx.cpu().numpy()-dy64@w64.T).max()
    print(f"   dw err {e1:.3e} (rms {np.sqrt(((x64.T@dy64)**2).mean()):.2f})  dx err {e2:.3e} (rms {np.sqrt(((dy64@w64.T)**2).mean()):.2f})")
