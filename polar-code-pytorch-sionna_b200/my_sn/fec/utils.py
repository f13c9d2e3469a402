import torch as tc


def int_mod_2(x):
  """x mod 2 for integer-valued tensors (my_sn/fec/utils.py:2-13)."""
  return tc.bitwise_and(x.to(dtype=tc.int32), 1).to(dtype=x.dtype)
