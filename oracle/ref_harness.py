"""Import the UNMODIFIED reference from /root/reference (build container only; the GPU box has no
/root/reference).  Test infrastructure: used by oracle/gen_golden.py to produce tests/golden/*.npz and
to validate oracle/polar_oracle.py.  Never imported by the product or by tests that run on the GPU box.

Shims follow SURVEY.md 8(c): the reference imports three packages that are absent here
(importlib_resources, matplotlib, pyrallis) at module scope only; none of them does arithmetic.
"""
import copy
import dataclasses
import importlib.resources
import os
import sys
import types

REF = os.environ.get("POLAR_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "x_run_sn_polar"))


def install():
    """Put the reference on sys.path with the three import-time shims.  Idempotent."""
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF)
    os.environ["PYTHONBREAKPOINT"] = "0"          # stray breakpoint()s: crc.py:94, dec.py:661
    sys.dont_write_bytecode = True                 # reference dir is read-only
    for p in (REF, os.path.join(REF, "x_run_sn_polar")):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("importlib_resources", importlib.resources)   # my_sn/fec/polar/utils.py:4
    if "matplotlib" not in sys.modules:                                  # mapping.py:3, plotting.py:1
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p
    if "pyrallis" not in sys.modules:                                    # config.py:1, main.py:18,42
        pyr = types.ModuleType("pyrallis")

        def field(default=None, is_mutable=False, **kw):
            if is_mutable:
                return dataclasses.field(default_factory=lambda: copy.deepcopy(default))
            return dataclasses.field(default=default)

        pyr.field = field
        pyr.wrap = lambda *a, **k: (lambda f: f)
        sys.modules["pyrallis"] = pyr
    from my_sn.fec.crc import CRCEncoder
    CRCEncoder.device = "cpu"                      # crc.py:81 reads an attribute that is never set
