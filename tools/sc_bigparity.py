"""Full-batch parity of the GPU SC decoder against the C restatement of the reference (oracle/polar_oracle.c):
   python tools/sc_bigparity.py [n] [B] [ebno_db]      (GPU box; the oracle runs on all host threads)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from oracle import polar_oracle as po, c_oracle as co

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
ebno = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
k = n // 2
dev = torch.device("cuda", 0)
fp = po.rm_frozen_pos(n, n - k)
tables = dk.code_tables(fp, n, dev)
_, _, x = dk.awgn_frontend(tables, B, po.ebnodb2no(ebno, 2, k / n), 2024)
_, up = dk.sc_decode(x, tables, want_info=False, want_packed=True)
torch.cuda.synchronize()
t0 = time.time()
ref = co.sc_decode_full(x.cpu().numpy(), po.frozen_vec(fp, n))
t1 = time.time()
w = up.cpu().numpy().view(np.uint32)
got = np.unpackbits(w.view(np.uint8).reshape(B, -1), axis=-1, bitorder="little")[:, :n]
bad = int((got != ref).any(axis=1).sum())
print("SC n=%d B=%d Eb/N0=%.1f dB: %d codewords differ from the C restatement (oracle %.1f s)" % (n, B, ebno, bad, t1 - t0))
