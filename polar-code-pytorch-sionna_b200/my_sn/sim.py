"""Monte-Carlo BER/BLER loop with the reference call surface (my_sn/sim.py:4-140).

Differences that matter on B200: the bit/block error counts of an iteration are produced by one
fused kernel (`polar_count_errors_f32`) into a device counter pair and read back once per iteration
(the stop rules of sim.py:107-133 need them on the host); when `torch.distributed` is initialised the
four counters are combined with one 4 x int64 all-reduce per iteration so every rank takes the same
stop decisions.  `mc_fun` is then expected to simulate its shard (batch_size is the
per-rank batch).  When `mc_fun` is the fused AWGN link model, `sim_ber` hands over to `sim_ber_device`, which keeps
the whole loop -- including the stop rules -- on the device."""
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


def _device_loop_ok(mc_fun, soft_estimates, count_fn):
  """The on-device loop applies to link models that provide the `device_frontend` hook (fused AWGN / BEC front end)
  together with a decoder that has the packed fast path."""
  return (count_fn is None and not soft_estimates and getattr(mc_fun, "fused", False)
          and callable(getattr(mc_fun, "device_frontend", None))
          and not getattr(mc_fun, "cw_estimates", True) and hasattr(getattr(mc_fun, "decoder", None), "decode_packed")
          and tc.cuda.is_available())


class StopPredictor:
  """Should iteration ii+1 of an SNR point be queued before the stop flag of iteration ii has come back?

  The device loop never lets the GPU wait for the host inside a point: it queues one iteration ahead.  The price is one
  discarded iteration per point (the one queued past the stop) -- 50 % of the work of a sweep whose points stop after a
  single iteration (BASELINE configs[3]: 65536 block errors against a target of 1000).  So the look-ahead is skipped when
  the counters seen so far say, with margin, that the iterations already queued will reach a target: rate per iteration
  from this point's known iterations, else a tenth of the previous point's (error rates fall with Eb/N0).  Either
  misprediction only costs time -- a discarded iteration or one host round trip -- never a different result."""

  def __init__(self, target_bit_errs, target_block_errs, max_mc_iter, damp=0.1, first_point_prior=None):
    self.targets = (target_bit_errs, target_block_errs)
    self.max_iter = int(max_mc_iter)
    self.damp = damp
    # (bit errors, block errors) per iteration of the previous point; before the first point: the caller's prior (the
    # device loop passes a block error rate of 1 and a bit error rate of 1/2 -- sweeps start at their lowest Eb/N0 --
    # which `damp` turns into 0.1 / 0.05)
    self.prev_rate = first_point_prior
    self.start_point()

  def start_point(self):
    self.known_iters, self.known = 0, (0, 0)

  def observe(self, iters, bit_errs, block_errs):
    self.known_iters, self.known = int(iters), (int(bit_errs), int(block_errs))

  def end_point(self):
    if self.known_iters:
      self.prev_rate = tuple(c / self.known_iters for c in self.known)

  def speculate(self, queued):
    """queued: iterations of this point already queued (results of the last queued - known_iters unknown)."""
    if queued >= self.max_iter:
      return False
    if self.known_iters:
      rate = tuple(c / self.known_iters for c in self.known)
    elif self.prev_rate is not None:
      rate = tuple(c * self.damp for c in self.prev_rate)
    else:
      return True
    for tgt, cum, r in zip(self.targets, self.known, rate):
      if tgt is None:
        continue
      pred = cum + r * (queued - self.known_iters)
      if pred >= tgt + 3.0 * pred ** 0.5 + 1.0:
        return False                 # the queued iterations will (almost surely) reach the target: wait for the flag
    return True


def sim_ber_device(model, ebno_dbs, batch_size, max_mc_iter, target_bit_errs=None, target_block_errs=None,
                   early_stop=True, verbose=True, return_counters=False, lookahead=True, stats=None, profile=False):
  """the Monte-Carlo loop of sim.py:79-133 with every per-iteration step on the device.

  One iteration = front-end kernel (bits -> encoder -> QPSK -> AWGN -> logits; `model.device_frontend`) -> decoder kernel
  (bit-packed decisions) -> packed error counter -> [sharded: one 4 x int64 NCCL all-reduce] -> `polar_mc_control`, a
  one-thread kernel that accumulates the counters and evaluates the stop rules (target bit errors, target block errors,
  max iterations).  All buffers are allocated once per call.  The host never reads a counter synchronously inside an SNR
  point: it keeps ONE iteration queued ahead (unless `StopPredictor` says the queued work will reach a target) and polls
  the stop flag of the iteration before from pinned memory; an iteration queued past the stop is ignored by the control
  kernel and its random numbers are handed to the next point, which makes the result identical to the host loop
  (`sim_ber(..., on_device=False)`) for the same seed.
  Returns (ber, bler) like sim_ber; with return_counters also the int64 [P,4] counters, status and iterations.
  `stats` (dict, optional) receives {"queued": iterations launched, "counted": iterations counted}; with `profile`
  also "split_us": mean device time per iteration of front end / decoder / counter / all-reduce / control (CUDA events
  on the launching stream at the stage boundaries) and the host wall time per iteration."""
  dist = _dist()
  rank0 = dist is None or dist.get_rank() == 0
  verbose = verbose and rank0
  dev = dk.cuda_device(model.device)
  dec = model.decoder
  frozen_pos = getattr(model.encoder, "frozen_pos", None)
  if frozen_pos is None:
    frozen_pos = dec.frozen_pos
  tables = dk.code_tables(frozen_pos, model.n, dev)
  if model._seed is None:
    model._seed = int(tc.randint(0, 2 ** 62, (1,)).item())
  ebno_dbs = np.asarray(ebno_dbs, dtype=np.float32)
  P = ebno_dbs.shape[0]
  B = int(batch_size)
  max_mc_iter = int(max_mc_iter)
  counters = np.zeros((P, 4), dtype=np.int64)
  status = np.zeros(P, dtype=np.int64)
  iters = np.zeros(P, dtype=np.int64)
  runtime = np.zeros(P)
  n_queued = 0
  status_levels = ["not simulated", "reached max iter       ", "no errors - early stop",
                   "reached target bit errors", "reached target block errors"]
  fmt = "{: >9} |{: >11} |{: >11} |{: >12} |{: >12} |{: >13} |{: >12} |{: >12} |{: >10}"
  DEPTH = 2 if lookahead else 1
  world = dist.get_world_size() if dist is not None else 1
  pred = StopPredictor(target_bit_errs, target_block_errs, max_mc_iter,
                       first_point_prior=(0.5 * B * tables.k * world, 1.0 * B * world))
  with tc.cuda.device(dev):
    nw = dk.words(model.n)
    u_tx = tc.empty((B, nw), dtype=tc.int32, device=dev)
    llr = tc.empty((B, model.n), dtype=tc.float32, device=dev)
    u_hat = tc.empty((B, nw), dtype=tc.int32, device=dev)
    state = tc.zeros(8, dtype=tc.int64, device=dev)
    delta = tc.zeros(4, dtype=tc.int64, device=dev)
    sizes = tc.tensor([0, 0, B * tables.k, B], dtype=tc.int64, device=dev)
    host = [tc.zeros(8, dtype=tc.int64).pin_memory() for _ in range(DEPTH)]
    events = [tc.cuda.Event() for _ in range(DEPTH)]

    marks = []                                               # profile: 6 timing events per queued iteration

    def mark(row):
      if profile:
        e = tc.cuda.Event(enable_timing=True)
        e.record()
        row.append(e)

    def queue(i, ii, offset0):
      row = []
      mark(row)
      model.device_frontend(tables, B, ebno_dbs[i], model._seed, offset0 + ii * B, (u_tx, llr))
      mark(row)
      dec.decode_packed(llr, tables, out=u_hat)
      mark(row)
      delta.copy_(sizes)                                     # (0, 0, bits, blocks) of this rank's shard
      dk.count_errors_packed(u_tx, u_hat, tables.info_mask, model.n, delta)
      mark(row)
      if dist is not None:
        dist.all_reduce(delta)                               # stream-ordered NCCL all-reduce of 4 x int64
      mark(row)
      dk.mc_control(delta, state, target_bit_errs, target_block_errs, max_mc_iter)
      host[ii % DEPTH].copy_(state, non_blocking=True)
      events[ii % DEPTH].record()
      mark(row)
      if profile:
        marks.append(row)

    for i in range(P):
      t0 = time.perf_counter()
      state.zero_()
      offset0 = model._offset
      pred.start_point()
      queued = known = 0
      while True:
        if queued < max_mc_iter and queued - known < DEPTH and (queued == known or pred.speculate(queued)):
          queue(i, queued, offset0)
          queued += 1
          continue
        if known == queued:
          break
        events[known % DEPTH].synchronize()                  # oldest iteration whose outcome is still unknown
        h = host[known % DEPTH]
        known += 1
        pred.observe(int(h[6]), int(h[0]), int(h[1]))
        if int(h[4]):
          break
      n_queued += queued
      tc.cuda.current_stream(dev).synchronize()
      st = state.cpu().numpy()
      pred.observe(int(st[6]), int(st[0]), int(st[1]))
      pred.end_point()
      counters[i] = st[:4]; status[i] = st[5]; iters[i] = st[6]
      model._offset = offset0 + int(st[6]) * B                 # random numbers of a discarded iteration are reused
      runtime[i] = time.perf_counter() - t0
      if verbose:
        if i == 0:
          print(fmt.format("EbNo [dB]", "BER", "BLER", "bit errors", "num bits", "block errors", "num blocks",
                           "runtime [s]", "status")); print('-' * 135)
        ber_i = counters[i, 0] / counters[i, 2] if counters[i, 2] else 0.0
        bler_i = counters[i, 1] / counters[i, 3] if counters[i, 3] else 0.0
        print(fmt.format(str(np.round(ebno_dbs[i], 3)), f"{ber_i:.4e}", f"{bler_i:.4e}", int(counters[i, 0]),
                         int(counters[i, 2]), int(counters[i, 1]), int(counters[i, 3]), np.round(runtime[i], 1),
                         status_levels[int(status[i])]))
      if early_stop and counters[i, 1] == 0:                   # sim.py:128-133
        status[i] = 2
        if verbose:
          print(f"\nSimu stopped as no error occurred @ EbNo = {ebno_dbs[i]:.1f} dB.\n")
        break
  if stats is not None:
    stats.update(queued=int(n_queued), counted=int(iters.sum()), blocks=int(counters[:, 3].sum()))
    if profile and marks:
      tc.cuda.synchronize(dev)
      names = ("front_end", "decode", "count", "all_reduce", "control")
      split = {nm: float(np.mean([r[j].elapsed_time(r[j + 1]) for r in marks])) * 1e3 for j, nm in enumerate(names)}
      split["device_total"] = float(np.mean([r[0].elapsed_time(r[5]) for r in marks])) * 1e3
      split["host_wall"] = float(runtime.sum()) / max(len(marks), 1) * 1e6
      stats["split_us"] = split
  with np.errstate(divide='ignore', invalid='ignore'):
    ber = np.nan_to_num(counters[:, 0] / counters[:, 2])
    bler = np.nan_to_num(counters[:, 1] / counters[:, 3])
  out = (tc.from_numpy(ber.astype(np.float32)), tc.from_numpy(bler.astype(np.float32)))
  if return_counters:
    return out + (counters, status, iters)
  return out


def sim_ber(mc_fun, ebno_dbs, batch_size, max_mc_iter, soft_estimates=False, target_bit_errs=None,
            target_block_errs=None, early_stop=True, verbose=True, dtype=tc.complex64, device='cpu', count_fn=None,
            on_device=True):
  """Returns (ber, bler) per SNR point; same stop rules and status codes as sim.py:19-140.
  `count_fn(b, b_hat) -> (bit_errors, block_errors)` defaults to the CUDA counter kernel; it exists so the
  sharding / stop logic can be exercised with a test double on machines without a GPU."""
  if on_device and _device_loop_ok(mc_fun, soft_estimates, count_fn):
    return sim_ber_device(mc_fun, ebno_dbs, batch_size, max_mc_iter, target_bit_errs=target_bit_errs,
                          target_block_errs=target_block_errs, early_stop=early_stop, verbose=verbose)
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
