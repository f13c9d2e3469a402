"""Large-batch parity of the GPU list decoder against the C restatement of the reference (oracle/polar_oracle.c):
   python tools/scl_bigparity.py [n] [L] [B] [ebno_db]      (GPU box; the oracle runs on all host threads)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from oracle import polar_oracle as po, c_oracle as co

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
ebno = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
k = n // 2
dev = torch.device("cuda", 0)
fp = po.rm_frozen_pos(n, n - k)
tables = dk.code_tables(fp, n, dev)
_, _, x = dk.awgn_frontend(tables, B, po.ebnodb2no(ebno, 2, k / n), 4242)
r = dk.scl_decode(x, tables, L, want_packed=True, want_info=False, want_pm=True, want_list=True)
torch.cuda.synchronize()
t0 = time.time()
u_ref, pm_ref = co.scl_decode_full(x.cpu().numpy(), po.frozen_vec(fp, n), L)
t1 = time.time()
w = r["list"].cpu().numpy().view(np.uint32)
got = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, L, -1).astype(np.uint8)[..., :n]
pm = r["pm"].cpu().numpy()
best_bad = int((got[:, 0] != u_ref[:, 0]).any(axis=1).sum())
rel0 = np.abs(pm[:, 0] - pm_ref[:, 0]) / np.maximum(np.abs(pm_ref[:, 0]), 1e-30)
list_bad = 0
for b in range(B):
    list_bad += set(map(bytes, got[b])) != set(map(bytes, u_ref[b]))
pm_exact = int((pm.view(np.int64) == pm_ref.view(np.int64)).all(axis=1).sum())
bad_idx = np.nonzero((got[:, 0] != u_ref[:, 0]).any(axis=1))[0]
if len(bad_idx):
    # which side does the numpy restatement (the reference's own exp / log) take on those codewords?
    sub = x[torch.from_numpy(bad_idx[:8]).to(dev)].cpu().numpy()
    u_np, pm_np = po.scl_decode_full(sub, po.frozen_vec(fp, n), L)
    for j, bidx in enumerate(bad_idx[:8]):
        print("  codeword %d: numpy restatement agrees with the GPU: %s, with the C restatement: %s; best metrics GPU %.12g C %.12g numpy %.12g" %
              (bidx, np.array_equal(u_np[j, 0], got[bidx, 0]), np.array_equal(u_np[j, 0], u_ref[bidx, 0]), pm[bidx, 0], pm_ref[bidx, 0], pm_np[j, 0]))
print("n=%d L=%d B=%d Eb/N0=%.1f dB: best path differs on %d codewords, list (as a set) on %d, all L path metrics bit-identical on %d, "
      "max rel err of the best metric %.2e  (oracle %.1f s)" % (n, L, B, ebno, best_bad, list_bad, pm_exact, rel0.max(), t1 - t0))
