"""SC decoder: parity against the C restatement on a slice + timing + phase timeline.  python tools/sc_check.py [n ...]
(POLAR_SC3_DBG=1 adds the per-phase cycle counts of warp 0 of CTA 0; POLAR_SC_WARPS_SM=w overrides the warps per SM)"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from my_sn.trans.ebno import ebnodb2no
from oracle import polar_oracle as po, c_oracle as co
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
ns = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
def timeit(f, it=7):
    f(); f(); torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))
for n in ns:
    k = n // 2
    key = "rm_%d_%d" % (n, k)
    fp = fz[key] if key in fz else po.rm_frozen_pos(n, n - k)
    tables = dk.code_tables(fp, n, dev)
    B = (1 << 30) // n
    _, _, x = dk.awgn_frontend(tables, B, ebnodb2no(4.0 if n == 1024 else 3.0, 2, k / n), 1234)
    ui, got = dk.sc_decode(x, tables, want_info=True, want_packed=True)
    torch.cuda.synchronize()
    m = min(B, 4096)
    ref = co.sc_decode_full(x[B - m:].cpu().numpy(), po.frozen_vec(fp, n))
    gb = np.unpackbits(got[B - m:].cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[:, :n]
    bad = int((gb != ref).any(axis=1).sum())
    ok_info = torch.equal(ui, dk.unpack_info(got, tables.info_pos, n))
    out = torch.empty_like(got)
    t, med = timeit(lambda: dk.sc_decode(x, tables, want_info=False, out_packed=out))
    print("n=%d B=%d: %d of %d codewords differ from the oracle, info ok %s | %.3f ms (median %.3f)  %.3e cw/s  %.1f Gbit/s info  HBM-frac %.3f" %
          (n, B, bad, m, ok_info, t, med, B / t * 1e3, B / t * 1e3 * k / 1e9, B / t * 1e3 * (4 * n + k / 8) / 6552.3e9), flush=True)
    if os.environ.get("POLAR_SC3_DBG") == "1":
        L = ctypes.CDLL(dk.LIB_PATH); buf = (ctypes.c_ulonglong * 8)()
        rd = L.polar_sc5_debug_read if n >= 1024 else L.polar_sc4_debug_read
        rd(buf); dk.sc_decode(x, tables, want_info=False, out_packed=out); torch.cuda.synchronize(); rd(buf)
        v = list(buf); nb = max(v[7], 1)
        print("  warp0 cycles/batch: ch-descents %d  scr-descents %d  tmem-steps %d  bottom %d  merge %d  out %d | total %d batches %d" % tuple([a // nb for a in v[:7]] + [v[7]]))
