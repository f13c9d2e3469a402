"""Small fixed workloads for ncu captures (GPU box):  python tools/ncu_target.py sc|scl|sc2048|sc4096 [reps]
sc  = the bench's headline launch (SC k=512 n=1024, B=2^20, 4 dB); scl = configs[2] (L=8 + CRC11, B=2^18, 3 dB)."""
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

what = sys.argv[1] if len(sys.argv) > 1 else "sc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
if what.startswith("sc") and not what.startswith("scl"):
    n = {"sc": 1024, "sc2048": 2048, "sc4096": 4096, "sc512": 512}[what]
    k, B = n // 2, (1 << 30) // n
    tables = dk.code_tables(fz["rm_%d_%d" % (n, k)], n, dev)
    _, _, x = dk.awgn_frontend(tables, B, ebnodb2no(4.0 if n == 1024 else 3.0, 2, k / n), 1234)
    out = torch.empty((B, n // 32), dtype=torch.int32, device=dev)
    for _ in range(reps):
        dk.sc_decode(x, tables, want_info=False, out_packed=out)
elif what == "fe":                                          # AWGN front end, n = 1024, 2^20 codewords (4 GiB of logits)
    n, k, B = 1024, 512, 1 << 20
    tables = dk.code_tables(fz["rm_1024_512"], n, dev)
    for _ in range(reps):
        dk.awgn_frontend(tables, B, ebnodb2no(4.0, 2, k / n), 1234)
elif what == "osd":                                         # OSD n = 128, k = 64, t = 2 (2081 candidates per codeword)
    g = np.load(os.path.join(ROOT, "tests", "golden", "osd.npz"))
    n, k, t, B = 128, 64, 2, 1 << 14
    rows = torch.from_numpy(dk.pack_rows(g["gm_128_64_t1"])).to(dev)
    x = torch.randn((B, n), device=dev) * 3
    for _ in range(reps):
        dk.osd_decode(x, rows, n, k, t)
elif what == "scl32":                                       # configs[3]: SCL L=32 n=2048, two waves of the persistent grid
    n, k, L, B = 2048, 1024, 32, 4736
    tables = dk.code_tables(fz["rm_2048_1024"], n, dev)
    _, _, x = dk.awgn_frontend(tables, B, ebnodb2no(2.0, 2, k / n), 4321)
    out = torch.empty((B, n // 32), dtype=torch.int32, device=dev)
    for _ in range(reps):
        dk.scl_decode(x, tables, L, want_info=False, out_packed=out)
else:
    from my_sn.fec.crc import CRCEncoder
    n, k, L, B = 1024, 512, 8, 1 << 18
    tables = dk.code_tables(fz["rm_1024_512"], n, dev)
    chk = CRCEncoder("CRC11", k)
    rows = torch.from_numpy(chk.syndrome_rows(tables.info_pos_np, n).view(np.int32).copy()).to(dev)
    _, _, x = dk.awgn_frontend(tables, B, ebnodb2no(3.0, 2, k / n), 4321)
    out = torch.empty((B, n // 32), dtype=torch.int32, device=dev)
    for _ in range(reps):
        dk.scl_decode(x, tables, L, crc_rows=rows, crc_len=chk.crc_length, want_info=False, out_packed=out)
torch.cuda.synchronize()
print("done", what)
