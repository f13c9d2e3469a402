"""Packed error counter timing + check against torch (one B200).   python tools/count_probe.py"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import d_kernels as dk
dev = torch.device("cuda", 0)
for n, B in ((1024, 1 << 20), (128, 1 << 22), (4096, 1 << 18), (8192, 1 << 16), (64, 1 << 20)):
    nw = max(1, n // 32)
    a = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, nw), dtype=torch.int32, device=dev)
    b = a.clone()
    flip = torch.rand((B, nw), device=dev) < 0.01
    b[flip] ^= torch.randint(1, 2 ** 31 - 1, (int(flip.sum()),), dtype=torch.int32, device=dev)
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (nw,), dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    dk.count_errors_packed(a, b, mask, n, cnt)
    d = (a ^ b) & mask
    bits = sum(int(((d >> s) & 1).sum()) for s in range(32))
    blocks = int((d != 0).any(dim=1).sum())
    ok = cnt.tolist() == [bits, blocks]
    f = lambda: dk.count_errors_packed(a, b, mask, n, cnt)
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print("count_errors_packed n=%d B=%d: %s  %.3f ms  %.0f GB/s read" % (n, B, "ok" if ok else "MISMATCH %s vs %s" % (cnt.tolist(), [bits, blocks]), ms, 2 * B * nw * 4 / ms / 1e6), flush=True)
