import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "polar-code-pytorch-sionna_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
# the package mirrors the reference's import layout: `my_sn.*` from the package root and
# `polar.*`, `d_kernels`, `config`, `z_sys_model.*` from x_run_sn_polar/ (reference main.py:4-7)
for p in (ROOT, PKG, os.path.join(PKG, "x_run_sn_polar")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _clear_polar_options():
    """tests set tuning options with util.set_opt (polar_set_option); none may leak into the next test"""
    yield
    dk = sys.modules.get("d_kernels")
    if dk is not None and getattr(dk, "_lib", None) is not None:
        dk.clear_options()
