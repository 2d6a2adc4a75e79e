#!/usr/bin/env python
"""Build spimagine_b200/libspimcuda.so (and libspimfft.so) with nvcc for sm_100a (cross-compiles without a GPU),
and libspimtiff.so (host-side strip decoders of the frame reader) with gcc.

    python -m spimagine_b200.build [--force] [--verbose]

The library is built in-tree so that it travels with the source snapshot; nothing is JIT-compiled
at import time and there is no fallback if it is missing.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["spv_api.cu", "spv_mip.cu", "spv_mip_axis.cu", "spv_mip_smem.cu", "spv_iso.cu", "spv_bricks.cu", "spv_comp.cu", "spv_display.cu", "spv_filter.cu"]
HEADERS = ["spv_common.cuh", "spv_kernels.h"]
LIB = os.path.join(HERE, "libspimcuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--threads", "0",
    # no implicit contraction: the ray setup must round like the reference's fp32 expressions
    # (explicit fmaf where the fast path wants it), IEEE division and square root
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-O2,-pthread",
    "-shared", "-cudart", "static",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libspimcuda.so cannot be built (there is no CPU implementation)")
    return exe


# the spectrum processor lives in its own library: only it depends on cuFFT
FFT_LIB = os.path.join(HERE, "libspimfft.so")
FFT_SOURCES = ["spv_fft.cu"]
FFT_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
             "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-cudart", "static"]


def fft_up_to_date():
    if not os.path.exists(FFT_LIB):
        return False
    t = os.path.getmtime(FFT_LIB)
    deps = [os.path.join(CSRC, f) for f in FFT_SOURCES] + [os.path.join(INCLUDE, "spimfft.h"), __file__]
    return all(os.path.getmtime(d) <= t for d in deps)


def build_fft(force=False, verbose=False):
    if not force and fft_up_to_date():
        return FFT_LIB
    cmd = [nvcc()] + FFT_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", INCLUDE, "-o", FFT_LIB] + [os.path.join(CSRC, f) for f in FFT_SOURCES] + ["-lcufft"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return FFT_LIB


# strip decoders of the frame reader (host code only, gcc)
TIFF_LIB = os.path.join(HERE, "libspimtiff.so")
TIFF_SOURCES = ["tiff_codecs.c"]


def build_tiff(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in TIFF_SOURCES] + [os.path.join(INCLUDE, "spimtiff.h"), __file__]
    if not force and os.path.exists(TIFF_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(TIFF_LIB) for d in deps):
        return TIFF_LIB
    cmd = [shutil.which("gcc") or "gcc", "-O2", "-std=c11", "-fPIC", "-fvisibility=hidden", "-shared", "-Wall",
           "-I", INCLUDE, "-o", TIFF_LIB] + [os.path.join(CSRC, f) for f in TIFF_SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return TIFF_LIB


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "spimcuda.h"), __file__]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", INCLUDE, "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_fft(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_tiff(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
