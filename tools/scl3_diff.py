"""Debug aid: where scl3 and scl2 disagree, which one agrees with the C oracle?  (GPU box)"""
import os, sys
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
k = n // 2
dev = torch.device("cuda", 0)
fp = po.rm_frozen_pos(n, n - k)
tables = dk.code_tables(fp, n, dev)
_, _, x = dk.awgn_frontend(tables, B, po.ebnodb2no(3.0, 2, k / n), 1234)
dk.set_option("POLAR_SCL_MODE", int("1"))
r2 = dk.scl_decode(x, tables, L, want_packed=True, want_info=False, want_pm=True, want_list=True)
dk.set_option("POLAR_SCL_MODE", int("2"))
r3 = dk.scl_decode(x, tables, L, want_packed=True, want_info=False, want_pm=True, want_list=True)
torch.cuda.synchronize()
dl = (r3["list"] != r2["list"]).any(dim=2).any(dim=1)
dpm = ((r3["pm"] - r2["pm"]).abs() / r2["pm"].abs().clamp_min(1e-30)).max(dim=1).values
idx = torch.nonzero(dl | (dpm > 1e-9)).flatten().cpu().numpy()
print("differing codewords:", len(idx), idx[:20])
sub = idx[:24]
lg = x[torch.from_numpy(sub).to(dev)].cpu().numpy()
u_ref, pm_ref = co.scl_decode_full(lg, po.frozen_vec(fp, n), L)
u_np, pm_np = po.scl_decode_full(lg[:6], po.frozen_vec(fp, n), L)
def unpack(w):
    w = w.cpu().numpy().view(np.uint32)
    return ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[:-1] + (-1,)).astype(np.uint8)[..., :n]
l2 = unpack(r2["list"][torch.from_numpy(sub).to(dev)]); l3 = unpack(r3["list"][torch.from_numpy(sub).to(dev)])
p2 = r2["pm"][torch.from_numpy(sub).to(dev)].cpu().numpy(); p3 = r3["pm"][torch.from_numpy(sub).to(dev)].cpu().numpy()
for j, b in enumerate(sub):
    e2 = np.abs(p2[j] - pm_ref[j]).max(); e3 = np.abs(p3[j] - pm_ref[j]).max()
    s2 = set(map(bytes, l2[j])) == set(map(bytes, u_ref[j])); s3 = set(map(bytes, l3[j])) == set(map(bytes, u_ref[j]))
    b2 = np.array_equal(l2[j][0], u_ref[j][0]); b3 = np.array_equal(l3[j][0], u_ref[j][0])
    extra = ""
    if j < 6:
        extra = " | numpy-oracle vs C-oracle: list %s maxdpm %.2e" % (set(map(bytes, u_np[j])) == set(map(bytes, u_ref[j])), np.abs(pm_np[j] - pm_ref[j]).max())
    print("cw %6d: scl2 vs C-oracle: best %s list %s max|dpm| %.3e ; scl3: best %s list %s max|dpm| %.3e%s" % (b, b2, s2, e2, b3, s3, e3, extra))
    if j < 3:
        print("   pm_ref", pm_ref[j]); print("   pm_2  ", p2[j]); print("   pm_3  ", p3[j])
