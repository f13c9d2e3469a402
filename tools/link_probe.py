"""Per-call device time of System_AWGN_model.forward (the bench's `link` leg) with allocator statistics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from polar.enc import PolarEncoder
from polar.polar_sc import SC_Dec
from z_sys_model.awgn_model import System_AWGN_model
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
n, k, Bl = 1024, 512, 1 << 18
fp = fz["rm_1024_512"]
big = torch.empty((1 << 20, n), dtype=torch.float32, device=dev)          # the bench holds its 4 GiB batch meanwhile
model = System_AWGN_model(n, k, PolarEncoder(fp, n, None), SC_Dec(fp, n, device=dev), device=dev, seed=77)
for _ in range(4):
    model(Bl, 4.0)
torch.cuda.synchronize()
rows = []
for i in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st0 = torch.cuda.memory_stats(dev)
    t0 = time.perf_counter(); a.record(); bits, bits_hat = model(Bl, 4.0); b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); st1 = torch.cuda.memory_stats(dev)
    rows.append((a.elapsed_time(b), (t1 - t0) * 1e3, st1["num_device_alloc"] - st0["num_device_alloc"], st1["num_device_free"] - st0["num_device_free"]))
print("device ms / host ms / cudaMalloc / cudaFree per call:")
for r in rows:
    print("  %.3f  %.3f  %d  %d" % r)
print("reserved GiB", torch.cuda.memory_reserved(dev) / 2 ** 30)
