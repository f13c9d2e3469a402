"""sc5 (POLAR_SC_MODE=4) against the default SC mapping on the same logits + timing.  python tools/sc5_check.py [n ...]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from my_sn.trans.ebno import ebnodb2no
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
ns = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]
ref_mode = int(os.environ.get("REF_MODE", "3"))
def timeit(f, it=5):
    f(); f(); torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for n in ns:
    k = n // 2
    key = "rm_%d_%d" % (n, k)
    if key in fz:
        fp = fz[key]
    else:
        from oracle import polar_oracle as po
        fp = po.rm_frozen_pos(n, n - k)
    tables = dk.code_tables(fp, n, dev)
    for B in (37, 5000, (1 << 30) // n):
        _, _, x = dk.awgn_frontend(tables, B, ebnodb2no(4.0 if n == 1024 else 3.0, 2, k / n), 1234)
        dk.set_option("POLAR_SC_MODE", ref_mode)
        _, ref = dk.sc_decode(x, tables, want_info=False, want_packed=True)
        torch.cuda.synchronize()
        dk.set_option("POLAR_SC_MODE", 4)
        ui, got = dk.sc_decode(x, tables, want_info=True, want_packed=True)
        torch.cuda.synchronize()
        bad = int((got != ref).any(dim=1).sum())
        ok_info = torch.equal(ui, dk.unpack_info(got, tables.info_pos, n))
        print("n=%d B=%d: mismatching codewords %d, info ok %s" % (n, B, bad, ok_info), flush=True)
    out = torch.empty_like(got)
    for mode in (ref_mode, 4):
        dk.set_option("POLAR_SC_MODE", mode)
        t = timeit(lambda: dk.sc_decode(x, tables, want_info=False, out_packed=out))
        print("  mode %d: %.3f ms  %.3e cw/s  %.1f Gbit/s info  HBM-frac %.3f" % (mode, t, B / t * 1e3, B / t * 1e3 * k / 1e9, B / t * 1e3 * (4 * n + k / 8) / 6552.3e9), flush=True)
    if os.environ.get("POLAR_SC3_DBG") == "1":
        L = ctypes.CDLL(dk.LIB_PATH); buf = (ctypes.c_ulonglong * 8)()
        L.polar_sc5_debug_read(buf); dk.sc_decode(x, tables, want_info=False, out_packed=out); torch.cuda.synchronize(); L.polar_sc5_debug_read(buf)
        v = list(buf); nb = max(v[7], 1)
        print("  sc5 warp0 cycles/batch: descents %d  g %d  f %d  bottom %d  merge %d  out %d | total %d batches %d" % tuple([a // nb for a in v[:7]] + [v[7]]))
