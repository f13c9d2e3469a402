"""Frozen-set construction by the Reed-Muller (row weight) rule.
Mirrors get_Kern_frozen_bits / get_Kern_frozen_bits2 (x_run_sn_polar/polar/froze.py:4-30).  Host side,
once per code; not a kernel."""
import math

import numpy as np
import torch as tc


def _kron_power(kern, stages, kron, clone):
  m = clone(kern)
  for _ in range(stages - 1):
    m = kron(kern, m)
  return m


def get_Kern_frozen_bits(n, f_num, kern):
  '''G = kern^{(x) log_b n}; freeze the f_num rows of least weight.  Returns (G, G_weights, frozen_pos)
  with frozen_pos a sorted int64 torch tensor, as froze.py:13-14.

  Tie hazard: for k = n/2 and even log2 n the weight classes tie at the cut and the
  reference lets `torch.argsort` (unstable, CPU) pick the members.  The same op is applied to the same
  CPU fp32 values here, so the code is identical to the reference's on the same host.'''
  base = kern.shape[0]
  stages = int(math.log(n, base))
  assert base ** stages == n, f"{n=}, is not power of {base=}"
  G = _kron_power(kern, stages, tc.kron, tc.clone)
  G_weights = tc.sum(G, dim=1)
  order = tc.argsort(G_weights.detach().to("cpu", tc.float32))
  frozen_pos = tc.sort(order[:f_num])[0]
  return G, G_weights, frozen_pos


def get_Kern_frozen_bits2(n, f_num, kern):
  '''NumPy twin (froze.py:17-30): np.argsort tie order, may differ from the torch variant.'''
  base = kern.shape[0]
  stages = int(math.log(n, base))
  assert base ** stages == n, f"{n=}, is not power of {base=}"
  G = _kron_power(np.asarray(kern), stages, np.kron, np.copy)
  G_weights = np.sum(G, axis=1)
  frozen_pos = np.sort(np.argsort(G_weights)[:f_num])
  return G, G_weights, frozen_pos
