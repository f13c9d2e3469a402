import torch as tc
from torch import nn


class BinarySource(nn.Module):
  """Random 0/1 tensor of the requested shape (my_sn/trans/binary_source.py:18-19).  Plumbing: the
  fused link model draws its bits inside polar_awgn_frontend instead."""

  def __init__(self, dtype=tc.float32, device='cpu'):
    super().__init__()
    self.dtype = dtype
    self.device = device

  def forward(self, inputs):
    return tc.randint(0, 2, size=inputs, device=self.device, dtype=self.dtype)
