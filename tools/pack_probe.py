import os, sys, numpy as np, torch
ROOT=os.getcwd(); PKG=os.path.join(ROOT,"polar-code-pytorch-sionna_b200")
for p in (ROOT,PKG,os.path.join(PKG,"x_run_sn_polar")): sys.path.insert(0,p)
import d_kernels as dk
dev=torch.device("cuda",0)
for n,B in ((1024,1<<18),(128,1<<20),(64,1<<18),(4096,1<<16)):
    x=torch.randint(0,2,(B,n),device=dev).float()
    p=dk.pack_bits(x)
    ref=np.packbits(x.cpu().numpy().astype(np.uint8).reshape(B,-1,32),axis=2,bitorder="little").view(np.uint32).reshape(B,-1)
    ok=np.array_equal(p.cpu().numpy().view(np.uint32),ref)
    f=lambda: dk.pack_bits(x); f(); torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms=float(np.median(ts)); print("pack_bits n=%d B=%d: %s %.3f ms %.0f GB/s read"%(n,B,"ok" if ok else "MISMATCH",ms,B*n*4/ms/1e6))
