"""my_sn PolarEncoder (my_sn/fec/polar/enc.py:8-113): same transform as the x_run encoder, without a
G argument.  The gather/xor stages of G_matrix (:85-96) are the bit-packed butterfly kernel; the
reference's per-call parity self-check H.c = 0 (:110) is available as `check_parity`."""
import numpy as np
import torch as tc

from polar.enc import PolarEncoder as _Enc
import d_kernels as dk


class PolarEncoder(_Enc):
  def __init__(self, frozen_pos, n, dtype=tc.float32, device='cpu'):
    super().__init__(frozen_pos, n, None, dtype=dtype, device=device)
    self._nb_stages = int(np.log2(self._n))

  def check_parity(self, c):
    """H.c = 0 (mod 2): re-transforming a codeword must give zeros at every frozen position."""
    dev = c.device
    tables = dk.code_tables(self._frozen_pos, self._n, dev)
    u_back = dk.encode_packed(dk.pack_bits(c), self._n)              # G is an involution
    return bool(tc.all((u_back & tables.frozen_mask) == 0))
