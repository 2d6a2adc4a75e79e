#!/usr/bin/env python
"""Build the CPU oracle libraries.  TEST INFRASTRUCTURE ONLY (see spim_oracle.c).

  oracle/libspim_oracle.so      the C restatement (oracle/spim_oracle.c), always built
  oracle/_ref/libspim_ref.so    the REFERENCE's own kernel text compiled for the host,
                                built only where /root/reference exists (the dev
                                container); git-ignored, but it travels to the GPU box
  oracle/_ref/libspim_ref_fast.so   the same text with the host equivalents of the
                                reference's OpenCL build options (volumerender.py:
                                154-160: -cl-fast-relaxed-math, -cl-mad-enable, ...:
                                -O3 -ffast-math -mavx2 -mfma).  TIMING ONLY (bench.py's
                                reference arm and cpu_baseline): its pixels are not
                                used for parity, fast-math changes them

The reference build never copies reference sources into the repo: the .cl files are
read where they lie, passed through one syntax rewrite that OpenCL C needs to be valid
C++ ("(float4)(" vector literals -> "float4(" constructor calls; "#include<...>" of
sibling kernel files is dropped because the files are concatenated here instead), and
streamed to g++ on stdin between oracle/ocl_shim.hpp and oracle/ref_driver.inc.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_KERNELS = "/root/reference/spimagine/volumerender/kernels"
# concatenation order == include order of all_render_kernels.cl:24-26 and iso_kernel.cl:11-14
REF_FILES = ["utils.cl", "volume_kernel.cl", "convolve_2d.cl", "occlusion.cl", "iso_kernel.cl"]
CFLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
          "-fvisibility=hidden", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "spim_oracle.c")
    out = os.path.join(HERE, "libspim_oracle.so")
    if not force and _newer(out, [src]):
        return out
    subprocess.check_call(["gcc", "-std=gnu99"] + CFLAGS + ["-o", out, src, "-lm"])
    return out


def reference_translation_unit():
    parts = ['#include "ocl_shim.hpp"\n']
    for name in REF_FILES:
        with open(os.path.join(REF_KERNELS, name)) as f:
            text = f.read()
        text = re.sub(r"^\s*#include\s*<[a-z_0-9]+\.cl>\s*$", "", text, flags=re.M)
        text = text.replace("(float4)(", "float4(")
        parts.append('#line 1 "%s"\n' % name)
        parts.append(text)
        parts.append("\n")
    parts.append('#line 1 "ref_driver.inc"\n#include "ref_driver.inc"\n')
    return "".join(parts)


FAST_CFLAGS = ["-O3", "-fopenmp", "-ffast-math", "-mavx2", "-mfma", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
               "-Wno-unknown-pragmas", "-Wno-unused"]


def build_ref(force=False, fast=False):
    """Returns the path of libspim_ref.so (fast=True: libspim_ref_fast.so), or None when the reference tree is absent
    (then a previously built copy, if any, is used as is)."""
    outdir = os.path.join(HERE, "_ref")
    out = os.path.join(outdir, "libspim_ref_fast.so" if fast else "libspim_ref.so")
    if not os.path.isdir(REF_KERNELS):
        return out if os.path.exists(out) else None
    deps = [os.path.join(HERE, "ocl_shim.hpp"), os.path.join(HERE, "ref_driver.inc")] + \
           [os.path.join(REF_KERNELS, n) for n in REF_FILES]
    if not force and _newer(out, deps):
        return out
    os.makedirs(outdir, exist_ok=True)
    tu = reference_translation_unit()
    cmd = ["g++", "-std=gnu++17", "-x", "c++", "-"] + (FAST_CFLAGS if fast else CFLAGS) + \
          ["-Wno-narrowing", "-I", HERE, "-o", out, "-lm"]
    subprocess.run(cmd, input=tu.encode(), check=True)
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_oracle(force))
    print(build_ref(force))
    print(build_ref(force, fast=True))
