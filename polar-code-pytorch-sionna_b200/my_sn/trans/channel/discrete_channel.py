"""Binary erasure channel layer (my_sn/trans/channel/discrete_channel.py:5-107).
The reference samples the erasure pattern with a Gumbel-softmax pair (a Bernoulli(pb) draw made differentiable for
TensorFlow); here the pattern comes from the Philox stream of `polar_bec_llr` on the GPU.  Statistically equivalent,
never bitwise (like every random layer of this package)."""
import torch as tc
from torch import nn

import d_kernels as dk


class BinaryMemorylessChannel(nn.Module):
  def __init__(self, return_llrs=False, bipolar_input=False, llr_max=100., dtype=tc.float32, device='cpu'):
    super().__init__()
    self.dtype = dtype
    assert isinstance(return_llrs, bool), "return_llrs must be bool."
    self._return_llrs = return_llrs
    assert isinstance(bipolar_input, bool), "bipolar_input must be bool."
    self._bipolar_input = bipolar_input
    assert llr_max >= 0., "llr_max must be a positive scalar value."
    self._llr_max = tc.tensor(llr_max).to(dtype=self.dtype)
    signed = (tc.float16, tc.float32, tc.float64, tc.int8, tc.int16, tc.int32, tc.int64)
    if self._return_llrs:
      assert dtype in (tc.float16, tc.float32, tc.float64), "LLR outputs require non-integer dtypes."
    elif self._bipolar_input:
      assert dtype in signed, "Only, signed dtypes are supported for bipolar inputs."
    else:
      assert dtype in signed + (tc.uint8, tc.uint16, tc.uint32, tc.uint64), "Only, real-valued dtypes are supported."
    self._check_input = True
    self.device = device
    self._seed = None
    self._offset = 0

  @property
  def llr_max(self): return self._llr_max

  @llr_max.setter
  def llr_max(self, value):
    assert value >= 0, 'llr_max cannot be negative.'
    self._llr_max = tc.as_tensor(value).to(dtype=tc.float32)

  def _check_inputs(self, x):
    """Inputs must be binary (or bipolar); checked once like the reference (discrete_channel.py:39-50)."""
    if self._check_input:
      lo, hi = (-1, 1) if self._bipolar_input else (0, 1)
      assert bool(tc.all((x == lo) | (x == hi))), "Input must be binary."
      self._check_input = False


class BinaryErasureChannel(BinaryMemorylessChannel):
  def __init__(self, return_llrs=False, bipolar_input=False, llr_max=100., dtype=tc.float32, device='cpu'):
    super().__init__(return_llrs=return_llrs, bipolar_input=bipolar_input, llr_max=llr_max, dtype=dtype, device=device)
    assert dtype in (tc.float16, tc.float32, tc.float64, tc.int8, tc.int16, tc.int32, tc.int64), \
        "Unsigned integers are currently not supported."

  def forward(self, inputs):
    """[x, pb] -> LLRs (+-llr_max, 0 where erased) or ternary symbols (erasure = -1, or 0 for bipolar input)."""
    x, pb = inputs
    pb = float(min(max(float(pb), 0.), 1.))
    self._check_inputs(x)
    dev = x.device if x.is_cuda else dk.cuda_device(self.device)
    bits = x.to(device=dev, dtype=tc.float32)
    if self._bipolar_input:
      bits = (bits + 1) * 0.5
    assert bits.shape[-1] % 4 == 0, "last dimension must be a multiple of 4."
    if self._seed is None:
      self._seed = int(tc.randint(0, 2 ** 62, (1,)).item())
    rows = bits.numel() // bits.shape[-1]
    scale = float(self._llr_max) if self._return_llrs else 1.0
    y = dk.bec_llr(bits, pb, self._seed, self._offset, llr_max=scale)       # +-scale, 0 where erased
    self._offset += rows
    if not self._return_llrs:
      if self._bipolar_input:
        pass                                                                # -1 / +1, erasure indicator 0
      else:
        y = tc.where(y == 0, tc.full_like(y, -1.), (y + 1) * 0.5)           # 0 / 1, erasure indicator -1
    y = y.to(self.dtype)
    return y if x.is_cuda else y.to(x.device)
