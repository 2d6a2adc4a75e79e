#!/usr/bin/env python
"""What does the B200 texture unit compute for a trilinear fetch?  Compares tex3D on random volumes with
software models: fp32 weights, weights rounded / truncated to 8 or 9 fractional bits, result rounding."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from spimagine_b200 import VolumeRenderer


def model(vol, pos, wmode, bits):
    nz, ny, nx = vol.shape
    v = vol.astype(np.float64)
    u = pos[:, 0].astype(np.float32) * np.float32(nx) - np.float32(.5)
    w = pos[:, 1].astype(np.float32) * np.float32(ny) - np.float32(.5)
    t = pos[:, 2].astype(np.float32) * np.float32(nz) - np.float32(.5)
    out = 0
    fl = [np.floor(c).astype(np.int64) for c in (u, w, t)]
    fr = [(c - np.floor(c)).astype(np.float64) for c in (u, w, t)]
    if wmode == "round":
        fr = [np.floor(f * (1 << bits) + .5) / (1 << bits) for f in fr]
    elif wmode == "trunc":
        fr = [np.floor(f * (1 << bits)) / (1 << bits) for f in fr]
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                i = np.clip(fl[0] + dx, 0, nx - 1)
                j = np.clip(fl[1] + dy, 0, ny - 1)
                k = np.clip(fl[2] + dz, 0, nz - 1)
                wgt = (fr[0] if dx else 1 - fr[0]) * (fr[1] if dy else 1 - fr[1]) * (fr[2] if dz else 1 - fr[2])
                out = out + wgt * v[k, j, i]
    return out


rng = np.random.default_rng(0)
pos = rng.uniform(0.02, 0.98, (200000, 3)).astype(np.float32)
for dtype, scale in ((np.float32, 1.), (np.uint16, 65535.), (np.uint16, 255.), (np.uint8, 255.)):
    shape = (40, 48, 56)
    if dtype == np.float32:
        vol = rng.random(shape, dtype=np.float32)
    else:
        vol = rng.integers(0, int(scale) + 1, shape).astype(dtype)
    r = VolumeRenderer((8, 8))
    r.set_data(vol)
    hw = r.sample_points(pos).astype(np.float64)
    r.set_sampler("exact")
    ex = r.sample_points(pos).astype(np.float64)
    rng_ = float(vol.max()) - float(vol.min())
    print("%s data range %g:" % (np.dtype(dtype).name, rng_))
    print("   exact sampler vs fp64 model (fp32 weights): max %.3g" % (np.abs(ex - model(vol, pos, 'none', 0)).max()))
    for wmode, bits in (("none", 0), ("round", 8), ("trunc", 8), ("round", 9), ("trunc", 9), ("round", 7)):
        m = model(vol, pos, wmode, bits)
        d = np.abs(hw - m)
        print("   tmu vs model(weights %s %d bits): max %.4g  mean %.4g  (in units of the data range: max %.3g)" % (
            wmode, bits, d.max(), d.mean(), d.max() / rng_))
    m8 = model(vol, pos, "round", 8)
    if dtype != np.float32:
        full = 65535. if dtype == np.uint16 else 255.
        for qb in (8, 10, 12, 14, 16, 18, 20, 24):
            q = np.round(m8 / full * (1 << qb)) / (1 << qb) * full
            print("      + result rounded to %d bits of full scale: max %.4g mean %.4g" % (qb, np.abs(hw - q).max(), np.abs(hw - q).mean()))
    r.close()
