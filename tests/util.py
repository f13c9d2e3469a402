import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def unpack_words(w, n):
    """int32/uint32 words [..., nw] -> uint8 bits [..., n] (LSB first)."""
    w = np.ascontiguousarray(w).view(np.uint32)
    b = np.unpackbits(w.view(np.uint8).reshape(w.shape + (4,)), axis=-1, bitorder="little")
    return b.reshape(w.shape[:-1] + (-1,))[..., :n]


def pack_words(bits):
    """uint8 bits [..., n] -> uint32 words [..., max(1, n/32)]."""
    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    if n < 32:
        pad = np.zeros(bits.shape[:-1] + (32 - n,), dtype=np.uint8)
        bits = np.concatenate([bits, pad], axis=-1)
    return np.packbits(bits, axis=-1, bitorder="little").view(np.uint32)


def awgn_logits(rng, n, k, frozen_pos, B, ebno_db):
    """Synthetic BPSK/AWGN logits for random codewords of the given code (numpy, CPU)."""
    from oracle import polar_oracle as po
    u = rng.integers(0, 2, size=(B, k)).astype(np.uint8)
    c = po.encode(u, frozen_pos, n)
    no = po.ebnodb2no(ebno_db, 2, k / n)
    y = (1.0 - 2.0 * c) / np.sqrt(2.0) + np.sqrt(no / 2.0) * rng.standard_normal((B, n))
    return u, (-2.0 * np.sqrt(2.0) * y / no).astype(np.float32)


def set_opt(name, value):
    """Override a POLAR_* tuning option for the current test (polar_set_option; undone by conftest's autouse fixture)."""
    import d_kernels as dk
    dk.set_option(name, int(value))
