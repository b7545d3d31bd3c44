"""TEST INFRASTRUCTURE -- numpy/ctypes front-end of oracle/lgs_oracle_adam.c: one Adam step with the rounding sequence of
torch.optim.Adam's CUDA foreach path (the reference's optimizer, scene/gaussian_model.py:390).  Importable only from tests/
and __graft_entry__.smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liblgs_oracle_adam.so")
_SRC = os.path.join(HERE, "lgs_oracle_adam.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        fma = []
        try:
            if " fma " in open("/proc/cpuinfo").read():
                fma = ["-mfma"]
        except OSError:
            pass
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared"] + fma + ["-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p = C.POINTER(C.c_float)
        L.lgs_adam_oracle.restype = None
        L.lgs_adam_oracle.argtypes = [C.c_longlong, f32p, f32p, f32p, f32p] + [C.c_float] * 6 + [C.c_int]
        _lib = L
    return _lib


def scalars(lr, beta1, beta2, eps, step):
    """torch/optim/adam.py:773-781, in double"""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    return (1 - beta1, beta2, 1 - beta2, eps, (lr / bc1) * -1, bc2 ** 0.5)


def adam_step(p, g, m, v, lr, betas, eps, step, variant=0):
    """In place on float32 arrays p, m, v (any shape, contiguous); `step` is the 1-based step number."""
    f32p = C.POINTER(C.c_float)
    for a in (p, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    g = np.ascontiguousarray(g, np.float32)
    lib().lgs_adam_oracle(p.size, p.ctypes.data_as(f32p), g.ctypes.data_as(f32p), m.ctypes.data_as(f32p), v.ctypes.data_as(f32p),
                          *scalars(lr, betas[0], betas[1], eps, step), variant)
