"""Frozen-set helpers (my_sn/fec/polar/utils.py:6-101): 5G reliability ranking and RM codes.  Host side."""
import os

import numpy as np

_SEQ = None


def _polar_sequence():
  """Q_0..Q_1023 of TS 38.212 Tab. 5.3.1.2-1 in ascending reliability (compact form of the reference's
  codes/polar_5G.csv)."""
  global _SEQ
  if _SEQ is None:
    _SEQ = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "codes", "polar_5g_seq.npy")).astype(int)
  return _SEQ


def generate_5g_ranking(k, n, sort=True, strict=True):
  """[frozen_pos, info_pos] of the 5G polar code (utils.py:6-71)."""
  if strict:
    assert k < 1025, "k cant > 1024."; assert n < 1025, "n cant > 1024."; assert n > 31, "n cant < 32."
    assert n >= k, "Invalid coderate (>1)."; assert np.log2(n) == int(np.log2(n)), "n must be a power of 2."
  seq = _polar_sequence()
  seq_n = seq[seq < n]                       # sub-sequence for length n keeps the reliability order
  frozen_pos = seq_n[:n - k].copy()
  info_pos = seq_n[n - k:].copy()
  if sort:
    info_pos = np.sort(info_pos); frozen_pos = np.sort(frozen_pos)
  return [frozen_pos.astype(int), info_pos.astype(int)]


def generate_rm_code(r, m):
  """[frozen_pos, info_pos, n, k, d_min] of the RM(r, m) code (utils.py:73-101): keep rows of weight >= 2^(m-r)."""
  assert r <= m, "order r cannot be larger than m."
  n = 2 ** m
  w = np.array([bin(i).count("1") for i in range(n)])
  info = np.nonzero(w >= m - r)[0]
  frozen = np.nonzero(w < m - r)[0]
  return [frozen.astype(int), info.astype(int), n, int(info.shape[0]), 2 ** (m - r)]
