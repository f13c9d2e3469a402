"""tests/golden/illcond_scl_1024_L8.npz: the one codeword of the round-1 large-batch SCL parity run (n=1024, k=512, L=8,
B=2^17, seed 4242, 3 dB: codeword 122257) whose best path differed between the GPU and the CPU restatements.
Input: gpurun_out/cw122257.npz, written on the GPU box by tests/test_gpu_fullsize.py::test_scl_known_ill_conditioned_codeword
(the front end is a pure function of (seed, codeword index), so the logits are reproducible).
  python tools/make_illcond_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = np.load(os.path.join(ROOT, "gpurun_out", "cw122257.npz"))
bad = np.nonzero((d["gpu_best"] != d["c_best"]).any(axis=1))[0]
assert list(bad) == [3], bad
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "illcond_scl_1024_L8.npz"),
                    logits=d["logits"][3:4], gpu_best=np.packbits(d["gpu_best"][3:4], axis=-1, bitorder="little"),
                    c_best=np.packbits(d["c_best"][3:4], axis=-1, bitorder="little"), gpu_pm=d["gpu_pm"][3:4], c_pm=d["c_pm"][3:4])
print("written")
