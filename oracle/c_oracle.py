"""ctypes loader for oracle/libpolar_oracle.so (C restatement; test infrastructure, NOT product)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpolar_oracle.so")
    src = os.path.join(_HERE, "polar_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpolar_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpolar_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().oracle_num_threads())


def sc_decode_full(logits, frozen, nthreads=0):
    """[B,n] fp32 logits, frozen uint8[n] -> u_hat uint8 [B,n] (all positions)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    frozen = np.ascontiguousarray(frozen, dtype=np.uint8)
    B, n = logits.shape
    out = np.zeros((B, n), dtype=np.uint8)
    lib().oracle_sc_decode(_p(logits), _p(frozen), ctypes.c_int(n), ctypes.c_long(B), _p(out), ctypes.c_int(nthreads))
    return out


def scl_decode_full(logits, frozen, L, nthreads=0):
    """-> (u_list uint8 [B,L,n] pm-ascending, pm float64 [B,L])."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    frozen = np.ascontiguousarray(frozen, dtype=np.uint8)
    B, n = logits.shape
    u = np.zeros((B, L, n), dtype=np.uint8)
    pm = np.zeros((B, L), dtype=np.float64)
    lib().oracle_scl_decode(_p(logits), _p(frozen), ctypes.c_int(n), ctypes.c_int(L), ctypes.c_long(B),
                            _p(u), _p(pm), ctypes.c_int(nthreads))
    return u, pm


def polar_transform(u_full):
    u_full = np.ascontiguousarray(u_full, dtype=np.uint8)
    B, n = u_full.shape
    c = np.zeros_like(u_full)
    lib().oracle_polar_transform(_p(u_full), ctypes.c_int(n), ctypes.c_long(B), _p(c))
    return c
