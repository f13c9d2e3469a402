"""CRC encoder / decoder with the reference call surface (my_sn/fec/crc.py:6-138), 38.212 polynomials.
The generator matrix is built on the host by LFSR stepping (crc.py:54-74); `syndrome_rows` exports it
as one 32-bit word per codeword position for the fused CRC-aided epilogue of the SCL kernel."""
import numpy as np
import torch as tc
from torch import nn

_POLYS = {  # TS 38.212 Sec. 5.1 (crc.py:40-45): exponents with non-zero coefficient
  "CRC24A": (24, [24, 23, 18, 17, 14, 11, 10, 7, 6, 5, 4, 3, 1, 0]),
  "CRC24B": (24, [24, 23, 6, 5, 1, 0]),
  "CRC24C": (24, [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]),
  "CRC16": (16, [16, 12, 5, 0]),
  "CRC11": (11, [11, 10, 9, 5, 0]),
  "CRC6": (6, [6, 5, 0]),
}


def crc_generator_words(crc_degree, k):
  """uint32[k]: word t = parity contribution of input bit t (bit (len-1-j) of the word = parity bit j)."""
  if crc_degree not in _POLYS:
    raise ValueError("Invalid CRC Polynomial")
  ln, exps = _POLYS[crc_degree]
  g = 0
  for e in exps:
    g |= 1 << e
  rows = np.zeros(k, dtype=np.uint32)
  rem = 1 << (ln - 1)                      # x^(len-1); one more shift gives x^len mod g for the last bit
  for t in range(k - 1, -1, -1):
    rem <<= 1
    if (rem >> ln) & 1:
      rem ^= g
    rows[t] = rem
  return rows, ln


class CRCEncoder(nn.Module):
  def __init__(self, crc_degree, k, dtype=tc.float32, device=None):
    super().__init__()
    assert isinstance(crc_degree, str), "crc_degree must be str"
    self.dtype = dtype
    self._crc_degree = crc_degree
    if crc_degree not in _POLYS:
      raise ValueError("Invalid CRC Polynomial")
    self._crc_length = _POLYS[crc_degree][0]
    pol = np.zeros(self._crc_length + 1, dtype=int)
    for e in _POLYS[crc_degree][1]:
      pol[self._crc_length - e] = 1          # MSB first (crc.py:48-52)
    self._crc_pol = pol
    self._k = k; self._n = None
    self.device = device
    self.build([None, k])

  @property
  def crc_degree(self): return self._crc_degree
  @property
  def crc_length(self): return self._crc_length
  @property
  def crc_pol(self): return self._crc_pol
  @property
  def k(self): return self._k
  @property
  def n(self): return self._n

  def build(self, input_shape):
    k = input_shape[-1]
    assert k is not None, "Shape of last dimension cannot be None."
    rows, ln = crc_generator_words(self._crc_degree, k)
    self._rows = rows
    shifts = np.arange(ln - 1, -1, -1, dtype=np.uint32)
    g = ((rows[:, None] >> shifts[None, :]) & 1).astype(np.float32)       # [k, len], crc.py:54-74
    self._g_mat_crc = tc.from_numpy(g)
    self._k = k
    self._n = k + ln

  def forward(self, inputs):
    """[...,k] -> [...,k+crc_length] (crc.py:84-109).  Runs on the input's device."""
    assert len(inputs.shape) > 1
    if inputs.shape[-1] != self._g_mat_crc.shape[0]:
      self.build(inputs.shape)
    g = self._g_mat_crc.to(inputs.device)
    x32 = inputs.to(dtype=tc.float32)
    par = tc.bitwise_and((x32 @ g).to(tc.int32), 1).to(dtype=self.dtype)
    return tc.concat([inputs.to(self.dtype), par], -1)

  def syndrome_rows(self, info_pos, n):
    """uint32[n] for polar_scl_decode: rows[info_pos[t]] = generator word of bit t, 0 at frozen positions
    (the CRC spans ALL k decoder outputs, payload and parity: crc.py:129-135, dec.py:508-516)."""
    assert len(info_pos) == self._k
    out = np.zeros(n, dtype=np.uint32)
    out[np.asarray(info_pos)] = self._rows
    return out


class CRCDecoder(nn.Module):
  def __init__(self, crc_encoder, dtype=tc.float32):
    super().__init__()
    assert isinstance(crc_encoder, CRCEncoder), "crc_encoder must be an instance of CRCEncoder."
    self._encoder = crc_encoder

  def forward(self, inputs):
    """(x, crc_valid) for inputs [...,k+crc_length] (crc.py:119-138): re-encodes all bits with the
    [k, len] generator matrix built for the full length; valid iff every parity output is 0."""
    if not isinstance(inputs, tc.Tensor):
      inputs = tc.from_numpy(np.asarray(inputs))
    assert len(inputs.shape) >= 2, "Input tensor must have at least rank 2."
    ln = self._encoder.crc_length
    assert inputs.shape[-1] >= ln, f"Last dimension of inputs must be at least {ln}."
    x_info = inputs[..., :-ln]
    x_parity = self._encoder(inputs)[..., -ln:]
    crc_check = tc.sum(x_parity, dim=-1, keepdim=True) <= 0
    return x_info, crc_check
