#!/usr/bin/env python
"""Static evidence that needs no GPU: per-kernel registers / shared memory / spills (ptxas -v) and a SASS mnemonic
census (cuobjdump -sass) of libspimcuda.so -> profiles/rNN_static_resources.txt, profiles/rNN_sass_census.txt.

    python scripts/static_census.py r01
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spimagine_b200 import build as b  # noqa: E402


def base(name):
    m = re.match(r"(?:spv::)?(?:\(anonymous namespace\)::)?([A-Za-z0-9_]+)", name.replace("void ", ""))
    return m.group(1) if m else name


def demangle(names):
    return subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()


def resources(tag):
    with tempfile.TemporaryDirectory() as tmp:      # the in-tree library is not touched
        cmd = [b.nvcc()] + b.NVCC_FLAGS + ["-Xptxas", "-v", "-I", b.INCLUDE, "-o", os.path.join(tmp, "x.so")] + \
              [os.path.join(b.CSRC, f) for f in b.SOURCES]
        txt = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for [^\n]+\n)?"
                      r"\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                      r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?([^\n]*)", txt)
    groups = collections.defaultdict(list)
    for e, n in zip(ents, demangle([e[0] for e in ents])):
        smem = re.search(r"(\d+) bytes smem", e[6])
        groups[base(n)].append((int(e[4]), int(smem.group(1)) if smem else 0, int(e[1]), int(e[2]), int(e[3])))

    def span(v):
        return "%d-%d" % (min(v), max(v)) if min(v) != max(v) else str(v[0])

    out = ["Static resources of every kernel in libspimcuda.so (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false,",
           "ptxas -v; one line per kernel template, over all its instantiations): the flags of spimagine_b200/build.py plus",
           "-Xptxas -v, compiled to a scratch file.  scripts/static_census.py", "",
           "%-28s %5s %12s %14s %12s %s" % ("kernel", "inst.", "registers", "static smem B", "stack B", "spill st/ld B (max)")]
    for k in sorted(groups):
        rs = groups[k]
        out.append("%-28s %5d %12s %14s %12s %d / %d" % (k, len(rs), span([r[0] for r in rs]), span([r[1] for r in rs]),
                                                        span([r[2] for r in rs]), max(r[3] for r in rs), max(r[4] for r in rs)))
    out += ["", "kernels with register spills: %s" % (sorted(k for k, rs in groups.items() if any(r[3] or r[4] for r in rs)) or "none")]
    open(os.path.join(ROOT, "profiles", "%s_static_resources.txt" % tag), "w").write("\n".join(out) + "\n")


def census(tag):
    txt = subprocess.run(["cuobjdump", "-sass", b.LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    cols = ["TEX", "TLD", "UTMALDG", "SYNCS", "LDG.E.128", "STG.E.128", "STG.E.64", "LDS", "STS", "SHFL", "VOTE", "ATOM", "RED",
            "BAR", "FFMA2", "FFMA", "FMNMX", "MUFU"]
    stats, count = collections.defaultdict(collections.Counter), collections.Counter()
    for f, n in zip(funcs, demangle([f.split("\n", 1)[0].strip() for f in funcs])):
        k = base(n)
        count[k] += 1
        for op in re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f):
            stats[k]["_total"] += 1
            for c in cols:
                if op == c or op.startswith(c + ".") or (c in ("TEX", "TLD") and op.startswith(c)):
                    stats[k][c] += 1
                    break
    out = ["SASS mnemonic census of spimagine_b200/libspimcuda.so (cuobjdump -sass, sm_100a), summed over the instantiations of",
           "each kernel template: what the kernels are made of.  TEX = filtered texture fetches (the hardware trilinear /",
           "bilinear sampler), TLD = unfiltered texel loads, UTMALDG = TMA tensor loads (cp.async.bulk.tensor), SYNCS = mbarrier",
           "operations (init / arrive / expect_tx / try_wait), STG.E.128 = 128-bit stores, SHFL / VOTE = warp-level exchange,",
           "FFMA2 = packed two-wide float32 fused multiply-add (fma.rn.f32x2 of sm_100).",
           "scripts/static_census.py", "",
           "%-28s %5s %8s " % ("kernel", "inst.", "SASS") + " ".join("%9s" % c for c in cols)]
    for k in sorted(stats):
        out.append("%-28s %5d %8d " % (k, count[k], stats[k]["_total"]) + " ".join("%9d" % stats[k][c] for c in cols))
    open(os.path.join(ROOT, "profiles", "%s_sass_census.txt" % tag), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    resources(tag)
    census(tag)
