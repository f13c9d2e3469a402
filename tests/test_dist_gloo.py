"""world_size-2 gloo test (CPU) of the sharded Monte-Carlo loop: two ranks each simulate half of every
batch; after the 4 x int64 all-reduce both must hold the single-process counters and take identical stop
decisions.  The decoder is replaced by the oracle here (CPU test double) -- the thing under
test is the host-side sharding / reduction / stop logic of my_sn.sim.sim_ber."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mc_fun_factory(rank, world, n, k, fp):
    from oracle import polar_oracle as po, c_oracle as co

    def mc_fun(batch_size, ebno_db):
        # shard `rank` of a global batch of world*batch_size codewords, deterministic per (ebno, shard)
        rng = np.random.default_rng(int(round(float(ebno_db) * 100)) * 10 + 7)
        u = rng.integers(0, 2, size=(world * batch_size, k)).astype(np.uint8)
        no = po.ebnodb2no(float(ebno_db), 2, k / n)
        noise = rng.standard_normal((world * batch_size, n))
        sl = slice(rank * batch_size, (rank + 1) * batch_size)
        c = po.encode(u[sl], fp, n)
        y = (1.0 - 2.0 * c) / np.sqrt(2.0) + np.sqrt(no / 2.0) * noise[sl]
        logits = (-2.0 * np.sqrt(2.0) * y / no).astype(np.float32)
        hat = co.sc_decode_full(logits, po.frozen_vec(fp, n))[:, po.info_positions(fp, n)]
        return torch.from_numpy(u[sl].astype(np.float32)), torch.from_numpy(hat.astype(np.float32))
    return mc_fun


def _count(b, b_hat):
    from oracle import polar_oracle as po
    return po.count_errors(b.numpy(), b_hat.numpy()), po.count_block_errors(b.numpy(), b_hat.numpy())


def _worker(rank, world, port, out):
    pkg = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
    for p in (ROOT, pkg, os.path.join(pkg, "x_run_sn_polar"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from my_sn.sim import sim_ber
    from oracle import polar_oracle as po
    n, k = 64, 32
    fp = po.rm_frozen_pos(n, n - k)
    ber, bler = sim_ber(_mc_fun_factory(rank, world, n, k, fp), np.arange(0, 8, 1.0), 50, 3, target_block_errs=40,
                        verbose=False, count_fn=_count)
    out[rank] = (ber.numpy().copy(), bler.numpy().copy())
    dist.destroy_process_group()


def test_sharded_sim_ber_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single process: same global batches (world=1 sees all world*50 codewords through 2 shards' union)
    pkg = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
    for p in (pkg, os.path.join(pkg, "x_run_sn_polar")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from my_sn.sim import sim_ber
    from oracle import polar_oracle as po
    n, k = 64, 32
    fp = po.rm_frozen_pos(n, n - k)
    f0, f1 = _mc_fun_factory(0, 2, n, k, fp), _mc_fun_factory(1, 2, n, k, fp)

    def both(batch_size, ebno_db):
        a, b = f0(batch_size // 2, ebno_db), f1(batch_size // 2, ebno_db)
        return torch.cat([a[0], b[0]]), torch.cat([a[1], b[1]])
    ber1, bler1 = sim_ber(both, np.arange(0, 8, 1.0), 100, 3, target_block_errs=40, verbose=False, count_fn=_count)
    for r in range(world):
        assert np.array_equal(out[r][0], ber1.numpy()) and np.array_equal(out[r][1], bler1.numpy())
    assert bler1[0] > bler1[3] and (bler1.numpy() >= 0).all()
    assert (ber1.numpy()[-1] == 0.0)                 # early stop leaves the tail at 0 (sim.py:128-139)


def test_sim_ber_stop_rules_single_process():
    from my_sn.sim import sim_ber
    calls = []

    def mc(batch_size, ebno_db):
        calls.append(float(ebno_db))
        b = torch.zeros(batch_size, 4)
        bh = b.clone()
        if ebno_db < 2:
            bh[: batch_size // 2, 0] = 1          # half the blocks wrong, one bit each
        return b, bh
    ber, bler = sim_ber(mc, np.array([0., 1., 2., 3.]), 10, 5, target_block_errs=12, verbose=False, count_fn=_count)
    # 5 block errors per call: target 12 reached after 3 iterations at 0 and 1 dB; 2 dB error-free -> early stop; 3 dB never run
    assert calls == [0.0] * 3 + [1.0] * 3 + [2.0] * 5
    assert np.allclose(bler.numpy(), [0.5, 0.5, 0.0, 0.0]) and np.allclose(ber.numpy(), [0.125, 0.125, 0, 0])
    ber, bler = sim_ber(mc, np.array([0., 1.]), 10, 4, target_bit_errs=6, early_stop=False, verbose=False, count_fn=_count)
    assert np.allclose(bler.numpy(), [0.5, 0.5])
