"""ctypes front-end of oracle/filter_oracle.c (CPU restatement of gputools.convolve_sep3).  TEST INFRASTRUCTURE ONLY:
see the header of filter_oracle.c for who may import this and for why the parity of this path is unpinned."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "filter_oracle.c")
LIB = os.path.join(HERE, "libspim_filter_oracle.so")
_FP = C.POINTER(C.c_float)
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(SRC) <= os.path.getmtime(LIB):
        return LIB
    # -ffp-contract=off: only the explicit fmaf calls fuse; -mfma lets gcc inline them where the CPU has FMA
    # (fmaf is correctly rounded either way)
    flags = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall"]
    try:
        with open("/proc/cpuinfo") as f:
            if " fma " in f.read():
                flags.append("-mfma")
    except OSError:
        pass
    subprocess.check_call(["gcc", "-std=gnu99"] + flags + ["-o", LIB, SRC, "-lm"])
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.sfo_convolve_sep3.restype = None
        lib.sfo_convolve_sep3.argtypes = [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP, C.c_int, _FP, C.c_int, _FP,
                                          C.c_int, C.c_int]
        _lib = lib
    return _lib


def convolve_sep3(data, hx, hy, hz, fused=True):
    """gputools.convolve_sep3(data, hx, hy, hz) for a numpy volume (Nz, Ny, Nx): float32 result."""
    lib = load()
    d = np.ascontiguousarray(np.asarray(data).astype(np.float32))
    hs = [np.ascontiguousarray(np.asarray(h, dtype=np.float64).astype(np.float32)) for h in (hx, hy, hz)]
    nz, ny, nx = d.shape
    res, tmp = np.empty_like(d), np.empty_like(d)
    p = lambda a: a.ctypes.data_as(_FP)  # noqa: E731
    lib.sfo_convolve_sep3(p(d), p(res), p(tmp), nx, ny, nz, p(hs[0]), len(hs[0]), p(hs[1]), len(hs[1]),
                          p(hs[2]), len(hs[2]), 1 if fused else 0)
    return res


def gauss_taps(sigma):
    """the taps BlurProcessor builds (spimagine/models/imageprocessor.py:52-55)"""
    N = 2 * sigma + 1
    x = np.arange(-N, N + 1)
    h = np.exp(-x ** 2 / 2. / sigma ** 2)
    return 1. * h / sum(h)
