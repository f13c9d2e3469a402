"""Polar encoder with the reference call surface (x_run_sn_polar/polar/enc.py:8-43).
The dense `(c @ G) % 2` of the reference (enc.py:42) is replaced by the bit-packed XOR butterfly
kernel `polar_encode_f32` (csrc/polar_enc.cu); both compute c_j = XOR_{i superset j} u_i.  `G` is
accepted for signature compatibility and CHECKED: only the Arikan kernel power F2^{(x) log2 n} is supported -- any
other matrix (the reference's d_kernels also ships non-Arikan kernels) raises instead of silently encoding Arikan."""
import numpy as np
import torch as tc
from torch import nn

import d_kernels as dk


def _assert_arikan(G, n):
  """G must be F2^{(x) log2 n}: G[i, j] = 1 iff j is a bitwise subset of i (froze.py:9-12).  One full compare per encoder."""
  Gc = tc.as_tensor(G).detach().to("cpu")
  assert tuple(Gc.shape) == (n, n), "G must be [n, n]."
  idx = np.arange(n, dtype=np.uint32)
  want = (idx[:, None] & idx[None, :]) == idx[None, :]
  assert np.array_equal(Gc.numpy() != 0, want) and bool(((Gc == 0) | (Gc == 1)).all()), \
    "polar_b200: only the Arikan kernel F2^(x)log2(n) is supported by the butterfly encoder (got another G)."


class PolarEncoder(nn.Module):
  def __init__(self, frozen_pos, n, G=None, dtype=tc.float32, device='cpu'):
    super().__init__()
    self.device = device
    self.dtype = dtype
    assert np.log2(n) == int(np.log2(n)), "n must be a power of 2."
    self._k = n - len(frozen_pos)
    self._n = n
    self._frozen_pos = frozen_pos
    self.info_pos = np.setdiff1d(np.arange(self._n), dk.to_numpy_pos(frozen_pos))
    assert self._k == len(self.info_pos), "invalid info_pos generated."
    if G is not None:
      _assert_arikan(G, n)
    self.G_ = None        # never stored: n x n fp32 is 64 MB at n = 4096 and the butterfly does not read it

  @property
  def k(self): return self._k

  @property
  def n(self): return self._n

  @property
  def frozen_pos(self): return self._frozen_pos

  def forward(self, u):
    """u [bs,k] (0/1) -> codewords [bs,n] of `dtype`, on the CUDA device (input device if CUDA)."""
    assert u.shape[-1] == self._k, "Last dim must be len k."
    dev = u.device if u.is_cuda else dk.cuda_device(self.device)
    tables = dk.code_tables(self._frozen_pos, self._n, dev)
    lead = list(u.shape[:-1])
    c = dk.encode_f32(u, tables)
    return c.reshape(lead + [self._n]).to(dtype=self.dtype)
