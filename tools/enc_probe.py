import os, sys, numpy as np, torch
ROOT=os.getcwd(); PKG=os.path.join(ROOT,"polar-code-pytorch-sionna_b200")
for p in (ROOT,PKG,os.path.join(PKG,"x_run_sn_polar")): sys.path.insert(0,p)
import d_kernels as dk
dev=torch.device("cuda",0)
fz=np.load("tests/golden/frozen_sets.npz")
for n,k,B in ((1024,512,1<<18),(128,64,1<<21),(4096,2048,1<<16)):
    tables=dk.code_tables(fz["rm_%d_%d"%(n,k)],n,dev)
    u=torch.randint(0,2,(B,k),device=dev).float()
    f=lambda: dk.encode_f32(u,tables)
    f(); torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms=float(np.median(ts)); print("encode_f32 n=%d B=%d: %.3f ms  %.3e cw/s  %.0f GB/s (4k+4n B per codeword)"%(n,B,ms,B/ms*1e3,B*(4*k+4*n)/ms/1e6))
