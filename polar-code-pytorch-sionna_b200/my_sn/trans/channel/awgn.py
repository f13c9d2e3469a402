"""AWGN channel layer (my_sn/trans/channel/awgn.py:19-29): y = x + sqrt(no) * CN(0,1)."""
import torch as tc
from torch import nn

from my_sn.utils import complex_normal, expand_to_rank


class AWGN(nn.Module):
  def __init__(self, device='cpu'):
    super().__init__()
    self._real_dtype = tc.float32
    self.device = device

  def forward(self, inputs):
    x, no = inputs
    noise = complex_normal(x.shape, device=x.device)
    no = expand_to_rank(tc.as_tensor(no, device=x.device), target_rank=len(x.shape), axis=-1)
    noise = noise * tc.sqrt(no.to(dtype=self._real_dtype)).to(dtype=noise.dtype)
    return x + noise
