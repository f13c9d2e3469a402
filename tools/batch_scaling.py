"""Strong-scaling emulation on one GPU: decode time of the per-rank batch at N = 1, 2, 4, 8 ranks for the two BLER-sweep
workloads (bench.py `sweep`), i.e. how much of an N-GPU speed-up the kernels' wave quantisation alone leaves.
   python tools/batch_scaling.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np
import torch
import d_kernels as dk
from my_sn.trans.ebno import ebnodb2no

dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))


def timeit(f, it=5):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for (name, n, k, L, total) in (("SCL-32 n=2048", 2048, 1024, 32, 1 << 16), ("SC n=1024", 1024, 512, 0, 1 << 20)):
    tables = dk.code_tables(fz["rm_%d_%d" % (n, k)], n, dev)
    _, _, x = dk.awgn_frontend(tables, total, ebnodb2no(2.0, 2, k / n), 1234)
    out = torch.empty((total, n // 32), dtype=torch.int32, device=dev)
    t1 = None
    for N in (1, 2, 4, 8):
        B = total // N
        if L:
            f = lambda: dk.scl_decode(x[:B], tables, L, want_info=False, out_packed=out[:B])
        else:
            f = lambda: dk.sc_decode(x[:B], tables, want_info=False, out_packed=out[:B])
        t = timeit(f, 3 if L else 7)
        t1 = t1 or t
        print("%s: per-rank batch %7d (N=%d): %8.3f ms -> decode-only speed-up %.2fx of %d" % (name, B, N, t, t1 / t, N), flush=True)
