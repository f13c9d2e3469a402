"""End-to-end BEC link model with the reference call surface (z_sys_model/bec_model.py:6-27).
`forward(batch_size, ebno_db)`: the second argument is the erasure probability (the reference reuses the name).
With `fused=True` BinarySource, PolarEncoder and the erasure channel are one launch of `polar_bec_frontend`."""
import torch as tc
from torch import nn

from my_sn.trans import binary_source
from my_sn.trans.channel.discrete_channel import BinaryErasureChannel
import d_kernels as dk


class System_BEC_model(nn.Module):
  def __init__(self, c, encoder, decoder, cw_estimates=False, device='cpu', fused=True, seed=None):
    super().__init__()
    self.cw_estimates = cw_estimates
    self.c = c
    self.n = c.n
    self.k = c.k
    self.coderate = self.k / self.n
    self.device = device
    self.fused = fused
    self.binary_src = binary_source.BinarySource(device=device)
    self.channel = BinaryErasureChannel(return_llrs=True, device=device)
    self.encoder = encoder; self.decoder = decoder
    self._seed = seed
    self._offset = 0

  def device_frontend(self, tables, batch_size, ebno_db, seed, offset, out):
    """Hook of the on-device Monte-Carlo loop (my_sn/sim.py::sim_ber_device): one launch of polar_bec_frontend into the
    caller-owned buffers out = (u_packed, logits); `ebno_db` is the erasure probability (bec_model.py:16-17)."""
    pe = float(min(max(float(ebno_db), 0.), 1.))
    dk.bec_frontend(tables, batch_size, pe, seed, offset, llr_max=float(self.channel.llr_max), out=out)

  def forward(self, batch_size, ebno_db):
    dev = dk.cuda_device(self.device)
    pe = float(min(max(float(ebno_db), 0.), 1.))
    if self.fused:
      frozen_pos = getattr(self.encoder, "frozen_pos", None)
      if frozen_pos is None:
        frozen_pos = self.decoder.frozen_pos
      tables = dk.code_tables(frozen_pos, self.n, dev)
      if self._seed is None:
        self._seed = int(tc.randint(0, 2 ** 62, (1,)).item())
      u_packed, c_packed, llr = dk.bec_frontend(tables, batch_size, pe, self._seed, self._offset,
                                                llr_max=float(self.channel.llr_max), want_codeword=self.cw_estimates)
      self._offset += int(batch_size)
      bits_hat = self.decoder(llr)
      if self.cw_estimates:
        pos = tc.arange(self.n, dtype=tc.int32, device=dev)
        return dk.unpack_info(c_packed, pos, self.n), bits_hat
      return dk.unpack_info(u_packed, tables.info_pos, self.n), bits_hat
    if self.binary_src.device != dev:
      self.binary_src = binary_source.BinarySource(device=dev)
    bits = self.binary_src([batch_size, self.k])
    codewords = self.encoder(bits)
    llr = self.channel([codewords, pe])
    bits_hat = self.decoder(llr)
    if self.cw_estimates:
      return codewords, bits_hat
    return bits, bits_hat
