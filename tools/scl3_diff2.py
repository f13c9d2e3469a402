import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch
import d_kernels as dk
from oracle import polar_oracle as po
n, L, B, k = 1024, 8, 32768, 512
dev = torch.device("cuda", 0)
fp = po.rm_frozen_pos(n, n - k)
tables = dk.code_tables(fp, n, dev)
_, _, x = dk.awgn_frontend(tables, B, po.ebnodb2no(3.0, 2, k / n), 1234)
sel = [296, 217]
lg = x[sel].cpu().numpy()
np.save(os.path.join(ROOT, "gpurun_out", "cw296_logits.npy"), lg)
os.environ["POLAR_SCL_MODE"] = "2"
r3 = dk.scl_decode(torch.from_numpy(lg).to(dev), tables, L, want_packed=True, want_info=False, want_pm=True, want_list=True)
print("scl3 alone on the 2 rows pm:", r3["pm"].cpu().numpy())
def sp(x, use_log1p=False):
    a = np.abs(x); sm = np.log(1 + np.exp(-a)); return np.where(np.signbit(x), a + sm, sm)
orig = po._softplus_neg
po._softplus_neg = sp
u_id, pm_id = po.scl_decode_full(lg, po.frozen_vec(fp, n), L)
po._softplus_neg = orig
u_o, pm_o = po.scl_decode_full(lg, po.frozen_vec(fp, n), L)
print("numpy oracle literal :", pm_o)
print("numpy oracle identity:", pm_id)
