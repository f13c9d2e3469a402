"""QAM constellation, Mapper, Demapper (my_sn/trans/mapping.py:7-241).  Only the 2-bit (QPSK = BPSK per
real dimension) constellation is on the hot path (awgn_model.py:23); it is what the sm_100a front end
(`polar_awgn_frontend`, `polar_qpsk_awgn_llr`) implements in closed form.  These layer classes keep the
reference's composable API; they work on whatever device their inputs live on."""
import numpy as np
import torch as tc
from torch import nn

from ..utils import expand_to_rank


def pam_gray(b):
  """Gray-labelled PAM point in {+-1, +-3, ...} for bit vector b (mapping.py:7-14)."""
  if len(b) > 1:
    return (1 - 2 * b[0]) * (2 ** len(b[1:]) - pam_gray(b[1:]))
  return 1 - 2 * b[0]


def qam(n_bits_per_sym, normalize=True):
  """2^n-point QAM, label of point i = binary repr of i, even bits -> real axis (mapping.py:15-48)."""
  assert n_bits_per_sym % 2 == 0 and n_bits_per_sym > 0
  c = np.zeros([2 ** n_bits_per_sym], dtype=np.complex64)
  for i in range(2 ** n_bits_per_sym):
    b = np.array(list(np.binary_repr(i, n_bits_per_sym)), dtype=np.int16)
    c[i] = pam_gray(b[0::2]) + 1j * pam_gray(b[1::2])
  if normalize:
    nd = n_bits_per_sym // 2
    c /= np.sqrt(1 / (2 ** (nd - 2)) * np.sum(np.linspace(1, 2 ** nd - 1, 2 ** (nd - 1)) ** 2))
  return c


class QamConstell(nn.Module):
  def __init__(self, n_bits_per_symbol, normalize=True, dtype=tc.complex64, device='cpu'):
    super().__init__()
    self.dtype = dtype
    self.device = device
    self.n_bits_per_sym = int(n_bits_per_symbol)
    self.normalize = normalize
    self.points = tc.from_numpy(qam(self.n_bits_per_sym, normalize)).to(dtype).to(device)

  def forward(self, x=None):
    return self.points


class Mapper(nn.Module):
  def __init__(self, constell=None, dtype=tc.complex64, device='cpu'):
    super().__init__()
    assert constell.dtype == dtype, "Constellation has wrong dtype."
    self.constell = constell
    self.device = self.constell.device

  def forward(self, inputs):
    nb = self.constell.n_bits_per_sym
    shape = [-1] + list(inputs.shape[1:-1]) + [inputs.shape[-1] // nb, nb]
    bits = inputs.reshape(shape).to(tc.int64)
    base = 2 ** tc.arange(nb - 1, -1, -1, device=inputs.device)
    return self.constell.points.to(inputs.device)[tc.sum(bits * base, dim=-1)]


class SymbolLogits2LLRs(nn.Module):
  """LLR_i = logsumexp over points with bit i = 1 minus logsumexp over points with bit i = 0 (mapping.py:151-205)."""

  def __init__(self, n_bits_per_sym):
    super().__init__()
    self.n_bits_per_sym = n_bits_per_sym
    npts = 2 ** n_bits_per_sym
    lab = np.array([[int(ch) for ch in np.binary_repr(i, n_bits_per_sym)] for i in range(npts)])
    self._c0 = tc.tensor(np.stack([np.where(lab[:, i] == 0)[0] for i in range(n_bits_per_sym)], axis=1), dtype=tc.int64)
    self._c1 = tc.tensor(np.stack([np.where(lab[:, i] == 1)[0] for i in range(n_bits_per_sym)], axis=1), dtype=tc.int64)

  def forward(self, inputs):
    c0, c1 = self._c0.to(inputs.device), self._c1.to(inputs.device)
    return tc.logsumexp(inputs[..., c1], dim=-2) - tc.logsumexp(inputs[..., c0], dim=-2)


class Demapper(nn.Module):
  def __init__(self, constell, dtype=tc.complex64):
    super().__init__()
    assert constell.dtype == dtype, "Constellation has wrong dtype."
    self.constell = constell
    self.device = self.constell.device
    self._logits2llrs = SymbolLogits2LLRs(self.constell.n_bits_per_sym)

  def forward(self, inputs):
    y, no = inputs
    pts = self.constell.points.to(y.device).reshape([1] * len(y.shape) + [-1])
    d2 = tc.abs(y.unsqueeze(dim=-1) - pts) ** 2
    no = tc.as_tensor(no, device=y.device)
    no = expand_to_rank(no, target_rank=len(d2.shape), axis=-1)
    llr = self._logits2llrs(-d2 / no)
    return llr.reshape(list(y.shape[:-1]) + [y.shape[-1] * self.constell.n_bits_per_sym])
