"""SC_Dec.forward on a PAGEABLE cpu tensor (what the reference's callers pass): staging threads x chunk size.
   python tools/host_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from polar.polar_sc import SC_Dec
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
n, k, B = 1024, 512, 1 << 18
fp = fz["rm_1024_512"]
dec = SC_Dec(fp, n, device=dev)
x = (torch.randn((B, n)) * 4).float()                      # pageable
if len(sys.argv) > 1 and sys.argv[1] == "pinned":
    B = 1 << 20
    x = (torch.randn((B, n)) * 4).float().pin_memory()
print("host cores:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
ref = None
for nt in ((4,) if x.is_pinned() else (4, 8, 12, 16)):
    for mb in ((16, 32, 64, 128, 256) if x.is_pinned() else (32, 64, 128)):
        dk.set_option("POLAR_HOST_COPY_THREADS", nt); dk.set_option("POLAR_HOST_CHUNK_MB", mb)
        out = dec(x); out = dec(x)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); out = dec(x); ts.append(time.perf_counter() - t0)
        if ref is None: ref = out.clone()
        assert torch.equal(out, ref)
        t = float(np.median(ts))
        print("threads %2d chunk %3d MB: %.2f ms  %.2f Gbit/s info  (%.1f GB/s of logits)" % (nt, mb, t * 1e3, B * k / t / 1e9, B * n * 4 / t / 1e9), flush=True)
