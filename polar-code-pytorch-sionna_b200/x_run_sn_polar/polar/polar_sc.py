"""SC decoder with the reference call surface (x_run_sn_polar/polar/polar_sc.py:5-133).
The Python recursion (_polar_decode_sc_tf, _cn_op_tf, _vn_op_tf, _decode_batch) is replaced by one
launch of the sm_100a kernel `polar_sc_decode_f32` (csrc/polar_sc.cu); decisions are bit-exact."""
import math

import numpy as np
import torch as tc
from torch import nn

import d_kernels as dk


class SC_Dec(nn.Module):
  """Min-sum successive-cancellation decoder.  Input logits [...,n] (ln P1/P0), output [...,k] hard bits."""

  def __init__(self, frozen_pos, n, output_dtype=tc.float32, device='cpu', mode='llr'):
    super().__init__()
    self.output_dtype = output_dtype
    self.n = n
    self.frozen_pos = frozen_pos
    self.k = self.n - len(self.frozen_pos)
    self.info_pos = np.setdiff1d(np.arange(self.n), dk.to_numpy_pos(frozen_pos))
    assert self.k == len(self.info_pos), "Internal error: invalid " "info_pos generated."
    self.llr_max = 30.
    if mode not in ("llr", "max"):
      raise Exception('error...')          # polar_sc.py:44-45
    self.mode = mode                       # both modes give min-sum (polar_sc.py:46 overrides the boxplus)
    self.kern_size = 2
    self._n_stages = int(math.log(n, self.kern_size))
    assert 2 ** self._n_stages == n, "n must be a power of 2."
    self.device = device
    self.complexity = None

  def decode_packed(self, logits, tables, out=None):
    """Device fast path of the on-device Monte-Carlo loop: logits [B,n] on the GPU -> bit-packed decisions
    int32 [B, n/32] (all n positions, frozen = 0), no fp32 [B,k] tensor is materialised."""
    return dk.sc_decode(logits, tables, want_info=False, want_packed=True, out_packed=out)[1]

  def forward(self, inputs):
    self.complexity = 0
    assert inputs.shape[-1] == self.n, "Last input dim must be of len n."
    assert len(inputs.shape) > 1
    dev = inputs.device if inputs.is_cuda else dk.cuda_device(self.device)
    tables = dk.code_tables(self.frozen_pos, self.n, dev)
    if inputs.is_cuda:
      u_hat, _ = dk.sc_decode(inputs, tables, want_info=True)
    else:         # CPU tensor in -> CPU tensor out (the reference's boundary): chunked, overlapped H2D / decode / D2H
      u_hat = dk.sc_decode_host(inputs, tables)
    output_shape = list(inputs.shape)
    output_shape[-1] = self.k
    output_shape[0] = -1
    return u_hat.reshape(output_shape).to(dtype=self.output_dtype)
