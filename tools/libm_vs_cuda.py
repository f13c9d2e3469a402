"""How often do CUDA's fp64 exp / log differ in the last bit from the host's (numpy) on the decoder's domain? (GPU box)"""
import numpy as np, torch
rng = np.random.default_rng(7)
N = 1 << 25
tot = {"exp": 0, "log1pexp_log": 0, "softplus": 0}
for rep in range(4):
    x = rng.uniform(-30, 30, N)
    xg = torch.from_numpy(x).cuda()
    e_g = torch.exp(xg).cpu().numpy(); e_h = np.exp(x)
    tot["exp"] += int((e_g.view(np.int64) != e_h.view(np.int64)).sum())
    w = 1.0 + e_h
    l_g = torch.log(torch.from_numpy(w).cuda()).cpu().numpy(); l_h = np.log(w)
    tot["log1pexp_log"] += int((l_g.view(np.int64) != l_h.view(np.int64)).sum())
    s_g = torch.log(1.0 + torch.exp(xg)).cpu().numpy(); s_h = np.log(1.0 + np.exp(x))
    tot["softplus"] += int((s_g.view(np.int64) != s_h.view(np.int64)).sum())
print("arguments:", 4 * N, "bitwise differences CUDA vs numpy:", tot)
