"""OSD decoder timing (CUDA events) on random polar-code words.   python tools/osd_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np
import torch
import d_kernels as dk

dev = torch.device("cuda", 0)
g = np.load(os.path.join(ROOT, "tests", "golden", "osd.npz"))
for (key, t, B) in (("64_32_t2", 2, 1 << 16), ("128_64_t1", 1, 1 << 16), ("128_64_t1", 2, 1 << 14), ("16_8_t3", 3, 1 << 18)):
    n, k = (int(v) for v in key.split("_")[:2])
    rows = torch.from_numpy(dk.pack_rows(g["gm_" + key])).to(dev)
    x = torch.randn((B, n), device=dev) * 3
    f = lambda: dk.osd_decode(x, rows, n, k, t)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    import math
    cands = sum(math.comb(k, w) for w in range(t + 1))
    print("OSD n=%d k=%d t=%d B=%d: %.2f ms  %.3e cw/s  (%d candidates per codeword, %.2e candidate-bits/s)" %
          (n, k, t, B, ms, B / ms * 1e3, cands, B / ms * 1e3 * cands * n), flush=True)
