"""Tensor helpers (my_sn/utils.py:2-58)."""
import torch as tc


def complex_normal(shape, var=1.0, device='cpu'):
  """CN(0, var): real and imaginary parts N(0, var/2), real drawn first (utils.py:2-17)."""
  std = tc.sqrt(tc.tensor(var / 2, dtype=tc.float32))
  xr = tc.normal(mean=0, std=std, size=shape, dtype=tc.float32, device=device)
  xi = tc.normal(mean=0, std=std, size=shape, dtype=tc.float32, device=device)
  return tc.complex(xr, xi)


def insert_dims(tensor, num_dims, axis=-1):
  assert num_dims >= 0, "`num_dims` must be nonnegative."
  rank = len(tensor.shape)
  assert -(rank + 1) <= axis <= rank, "`axis` is out of range `[-(D+1), D]`)"
  axis = axis if axis >= 0 else rank + axis + 1
  shape = tensor.shape
  return tensor.reshape(list(shape[:axis]) + [1] * num_dims + list(shape[axis:]))


def expand_to_rank(tensor, target_rank, axis=-1):
  return insert_dims(tensor, max(target_rank - len(tensor.shape), 0), axis)
