"""BASELINE configs[4]: mixed-length throughput table, SC and SCL L=4, n in {128,256,512,1024,2048,4096},
k = n/2, B = 2^28/n codewords per GPU (1 GiB of logits), Eb/N0 = 3 dB, CUDA-event timed, inputs resident in HBM.
   python tools/sweep_c5.py [out.md]                                   one GPU
   torchrun --nproc-per-node N ... tools/sweep_c5.py [out.md]          N GPUs: every rank decodes its own batch (weak scaling),
                                                                       barrier before each measurement, time = max over ranks"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import d_kernels as dk
from oracle import polar_oracle as po

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def timeit(fn, dev, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows = []
    for n in (128, 256, 512, 1024, 2048, 4096):
        k = n // 2
        B = (1 << 28) // n
        fp = po.rm_frozen_pos(n, n - k)
        tables = dk.code_tables(fp, n, dev)
        no = po.ebnodb2no(3.0, 2, k / n)
        _, _, x = dk.awgn_frontend(tables, B, no, 1234 + RANK)
        up = torch.empty((B, dk.words(n)), dtype=torch.int32, device=dev)
        f_sc = lambda: dk.check(dk.lib().polar_sc_decode_f32(dk.ptr(x), dk.ptr(tables.frozen_mask), n, B, dk.ptr(up), None, None, 0, dk.stream_ptr(dev)))
        ms_sc = timeit(f_sc, dev)
        Bs = min(B, 1 << 19)
        xs = x[:Bs]
        f_scl = lambda: dk.scl_decode(xs, tables, 4, want_packed=True, want_info=False)
        ms_scl = timeit(f_scl, dev, iters=3, warm=1)
        rows.append((n, k, B, WORLD * B / ms_sc * 1e3, WORLD * B / ms_sc * 1e3 * k / 1e9, B / ms_sc * 1e3 * (4 * n + k / 8) / 1e9,
                     Bs, WORLD * Bs / ms_scl * 1e3, WORLD * Bs / ms_scl * 1e3 * k / 1e9))
        if RANK == 0:
            print("n=%5d  SC %.3e cw/s %.1f Gbit/s (%.0f GB/s algorithmic per GPU)   SCL-4 %.3e cw/s %.2f Gbit/s   [%d GPU(s)]" %
                  (n, rows[-1][3], rows[-1][4], rows[-1][5], rows[-1][7], rows[-1][8], WORLD), flush=True)
        del x, up
    if RANK == 0 and len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write("# Mixed-length sweep (BASELINE configs[4] / SURVEY C5), %d B200, k = n/2 RM-rule code, Eb/N0 = 3 dB\n\n" % WORLD)
            f.write("`tools/sweep_c5.py` -- CUDA events, median of 5 (SC) / 3 (SCL) launches after warm-up, inputs resident in HBM; "
                    "batch per GPU, rates summed over the GPUs (time = max over ranks).\n\n")
            f.write("| n | k | SC batch / GPU | SC cw/s | SC info Gbit/s | SC algorithmic GB/s per GPU (4n+k/8) | SCL-4 batch / GPU | SCL-4 cw/s | SCL-4 info Gbit/s |\n|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %d | %d | %d | %.3e | %.1f | %.0f | %d | %.3e | %.2f |\n" % r)
        json.dump({"n_gpus": WORLD, "rows": rows}, open(sys.argv[1].replace(".md", ".json"), "w"))
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
