import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np, torch, time
import d_kernels as dk
from my_sn.trans.ebno import ebnodb2no
from my_sn.fec.crc import CRCEncoder
from my_sn.fec.polar.dec import SCL_Dec
dev = torch.device("cuda", 0)
fz = np.load(os.path.join(ROOT, "tests", "golden", "frozen_sets.npz"))
n, k, L, B = 1024, 512, 8, 1 << 18
fp = fz["rm_1024_512"]
tables = dk.code_tables(fp, n, dev)
chk = CRCEncoder("CRC11", k)
gen = CRCEncoder("CRC11", k - 11)
rows = torch.from_numpy(chk.syndrome_rows(tables.info_pos_np, n).view(np.int32).copy()).to(dev)
payload = torch.randint(0, 2, (B, k - 11), device=dev, dtype=torch.float32)
x = dk.qpsk_awgn_llr(dk.encode_f32(gen(payload), tables), ebnodb2no(3.0, 2, k / n), 4321)
a = dk.scl_decode(x, tables, L, crc_rows=rows, crc_len=11, want_info=True, want_packed=True, want_pm=True)
want = dk.unpack_info(a["u_packed"], tables.info_pos, n)
print("device info == unpack(packed):", torch.equal(a["u_info"], want))
b = dk.scl_decode(x, tables, L, crc_rows=rows, crc_len=11, want_info=False, want_packed=True)
print("packed-only == packed+info:", torch.equal(b["u_packed"], a["u_packed"]))
mod = SCL_Dec(fp, n, L, crc_degree="CRC11", cn_type="minsum", device=dev)
h = torch.empty((B, n), dtype=torch.float32, pin_memory=True); h.copy_(x)
for Bt in (2048, 40000, B):
    o = mod(h[:Bt])
    bad = (o != want[:Bt].cpu()).any(dim=1)
    print("host module B=%d: mismatching rows %d first %s pm ok %s" % (Bt, int(bad.sum()), bad.nonzero()[:5].flatten().tolist(),
          torch.equal(mod.msg_pm, a["pm"][:Bt].cpu())))
