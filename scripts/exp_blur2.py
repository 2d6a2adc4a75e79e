#!/usr/bin/env python
"""Blur passes with packed fma.rn.f32x2 (FFMA2): per-pass device times of every x / axis variant on an N^3 volume,
each result compared bit for bit with the scalar kernels' (x on single rows, one column per thread)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import scenes
from spimagine_b200 import imageprocessor as ip

N = int(os.environ.get("EXP_VOL", 512))
vf = ip.VolumeFilter(0)
base = scenes.vol_g(N, np.uint16, seed=0)


def run(t, vol, dt, taps, xp, ax, fuse=0, reps=6):
    vf.set_tuning(0, fuse); vf.set_tuning(1, ax); vf.set_tuning(2, xp)
    best, passes = 1e9, None
    for i in range(reps):
        vf.load_device(t.data_ptr(), vol.shape, dt)
        vf.convolve_sep3(*taps)
        vf.sync()
        if vf.last_ms() < best:
            best, passes = vf.last_ms(), vf.last_pass_ms()
    return best, passes


for dt in (np.uint16, np.float32, np.uint8):
    vol = base if dt == np.uint16 else (base.astype(np.float32) if dt == np.float32 else (base >> 8).astype(np.uint8))
    t = torch.from_numpy(vol.view(np.int16) if dt == np.uint16 else vol).cuda()
    for sigma in ((4., 1., 2., 7.) if dt == np.uint16 else (4.,)):
        taps = ip.BlurProcessor(sigma)._taps()
        ms, ps = run(t, vol, dt, taps, 0, 1)
        ref = vf.result()
        print("%-8s sigma %g (%2d taps)  scalar: %.3f ms  passes %s" % (np.dtype(dt).name, sigma, len(taps[0]), ms,
              " ".join("%.3f" % p for p in ps)), flush=True)
        for xp, ax in ((1, 1), (2, 1), (2, 1602), (2, 1604), (2, 804)):
            ms, ps = run(t, vol, dt, taps, xp, ax)
            same = bool(np.array_equal(vf.result(), ref))
            print("     x pairs %d, axis variant %4d: %.3f ms  passes %s  identical=%s" % (xp, ax, ms, " ".join("%.3f" % p for p in ps), same), flush=True)
        for xp, ax in ((2, 1604),):
            ms, ps = run(t, vol, dt, taps, xp, ax, fuse=1)
            same = bool(np.array_equal(vf.result(), ref))
            print("     fused x+y, axis variant %4d: %.3f ms  passes %s  identical=%s" % (ax, ms, " ".join("%.3f" % p for p in ps), same), flush=True)
    del t
# odd shapes: edges, rows that are not multiples of the vector width, fewer rows than a tile
for shape in ((70, 66, 130), (33, 1, 258), (5, 7, 3), (64, 65, 66), (200, 300, 1000), (3, 1000, 4)):
    rng = np.random.default_rng(3)
    vol = rng.integers(0, 60000, shape).astype(np.uint16)
    t = torch.from_numpy(vol.view(np.int16)).cuda()
    for sigma in (4., 1.):
        taps = ip.BlurProcessor(sigma)._taps()
        run(t, vol, np.uint16, taps, 0, 1, reps=1)
        ref = vf.result()
        ok = []
        for xp, ax in ((1, 1), (2, 1602), (2, 3202), (2, 1604), (2, 804)):
            run(t, vol, np.uint16, taps, xp, ax, reps=1)
            ok.append(bool(np.array_equal(vf.result(), ref)))
        print("shape", shape, "sigma", sigma, "identical:", ok, flush=True)
