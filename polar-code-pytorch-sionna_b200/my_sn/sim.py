"""Monte-Carlo BER/BLER loop with the reference call surface (my_sn/sim.py:4-140).

Differences that matter on B200: the bit/block error counts of an iteration are produced by one
fused kernel (`polar_count_errors_f32`) into a device counter pair and read back once per iteration
(the stop rules of sim.py:107-133 need them on the host); when `torch.distributed` is initialised the
four counters are combined with one 4 x int64 all-reduce per iteration so every rank takes the same
stop decisions.  `mc_fun` is then expected to simulate its shard (batch_size is the
per-rank batch)."""
import time

import numpy as np
import torch as tc

import d_kernels as dk


def hard_decisions(llr):
  return tc.where(llr > 0, 1., 0.)


def _count(b, b_hat):
  """(bit errors, block errors) between two 0/1 tensors (sim.py:7-18) as Python ints."""
  if b.is_cuda or b_hat.is_cuda:
    dev = b.device if b.is_cuda else b_hat.device
    counters = tc.zeros(2, dtype=tc.int64, device=dev)
    dk.count_errors_f32(b, b_hat, counters)
    c = counters.cpu()
    return int(c[0]), int(c[1])
  raise RuntimeError("polar_b200: count_errors needs CUDA tensors (no CPU fallback)")


def count_errors(b, b_hat):
  """Number of bit errors (sim.py:15-18); returns a 0-dim int64 tensor like the reference."""
  return tc.tensor(_count(b, b_hat)[0], dtype=tc.int64)


def count_block_errors(b, b_hat):
  """Number of rows with at least one differing element (sim.py:7-14)."""
  return tc.tensor(_count(b, b_hat)[1], dtype=tc.int64)


def _dist():
  import torch.distributed as dist
  return dist if (dist.is_available() and dist.is_initialized()) else None


def sim_ber(mc_fun, ebno_dbs, batch_size, max_mc_iter, soft_estimates=False, target_bit_errs=None,
            target_block_errs=None, early_stop=True, verbose=True, dtype=tc.complex64, device='cpu', count_fn=None):
  """Returns (ber, bler) per SNR point; same stop rules and status codes as sim.py:19-140.
  `count_fn(b, b_hat) -> (bit_errors, block_errors)` defaults to the CUDA counter kernel; it exists so the
  sharding / stop logic can be exercised with a test double on machines without a GPU."""
  count_fn = count_fn or _count
  dist = _dist()
  rank0 = dist is None or dist.get_rank() == 0
  verbose = verbose and rank0
  header = ["EbNo [dB]", "BER", "BLER", "bit errors", "num bits", "block errors", "num blocks", "runtime [s]", "status"]
  status_levels = ["not simulated", "reached max iter       ", "no errors - early stop",
                   "reached target bit errors", "reached target block errors"]
  fmt = "{: >9} |{: >11} |{: >11} |{: >12} |{: >12} |{: >13} |{: >12} |{: >12} |{: >10}"
  ebno_dbs = np.asarray(ebno_dbs, dtype=np.float32)
  num_points = ebno_dbs.shape[0]
  bit_errors = np.zeros(num_points, dtype=np.int64); block_errors = np.zeros(num_points, dtype=np.int64)
  nb_bits = np.zeros(num_points, dtype=np.int64); nb_blocks = np.zeros(num_points, dtype=np.int64)
  status = np.zeros(num_points, dtype=np.int64)
  runtime = np.zeros(num_points)

  def row(i, it, rt):
    ber = bit_errors[i] / nb_bits[i] if nb_bits[i] else 0.0
    bler = block_errors[i] / nb_blocks[i] if nb_blocks[i] else 0.0
    st = f"iter: {it:.0f}/{max_mc_iter:.0f}" if status[i] == 0 else status_levels[int(status[i])]
    return [str(np.round(ebno_dbs[i], 3)), f"{ber:.4e}", f"{bler:.4e}", int(bit_errors[i]), int(nb_bits[i]),
            int(block_errors[i]), int(nb_blocks[i]), np.round(rt, 1), st]

  for i in range(num_points):
    t0 = time.perf_counter()
    it = -1
    for ii in range(max_mc_iter):
      it += 1
      b, b_hat = mc_fun(batch_size=batch_size, ebno_db=ebno_dbs[i])[:2]
      if soft_estimates:
        b_hat = hard_decisions(b_hat)
      bit_e, block_e = count_fn(b, b_hat)
      bit_n = b.numel()
      block_n = int(b.numel() / b.shape[-1])
      if dist is not None:                       # 4 x int64 all-reduce: identical stop decisions on all ranks
        t = tc.tensor([bit_e, block_e, bit_n, block_n], dtype=tc.int64, device=b.device if b.is_cuda else 'cpu')
        dist.all_reduce(t)
        bit_e, block_e, bit_n, block_n = (int(v) for v in t.cpu())
      bit_errors[i] += bit_e; block_errors[i] += block_e
      nb_bits[i] += bit_n; nb_blocks[i] += block_n
      if verbose:
        if i == 0 and it == 0:
          print(fmt.format(*header)); print('-' * 135)
        print(fmt.format(*row(i, ii, time.perf_counter() - t0)), end="\r")
      if target_bit_errs is not None and bit_errors[i] >= target_bit_errs:
        status[i] = 3; break
      if target_block_errs is not None and block_errors[i] >= target_block_errs:
        status[i] = 4; break
      if it == max_mc_iter - 1:
        status[i] = 1
    runtime[i] = time.perf_counter() - t0
    if verbose:
      print(fmt.format(*row(i, it, runtime[i])))
    if early_stop and block_errors[i] == 0:
      status[i] = 2
      if verbose:
        print(f"\nSimu stopped as no error occurred @ EbNo = {ebno_dbs[i]:.1f} dB.\n")
      break
  with np.errstate(divide='ignore', invalid='ignore'):
    ber = np.nan_to_num(bit_errors / nb_bits)        # nan (never simulated) -> 0 (sim.py:134-139)
    bler = np.nan_to_num(block_errors / nb_blocks)
  return tc.from_numpy(ber.astype(np.float32)), tc.from_numpy(bler.astype(np.float32))
